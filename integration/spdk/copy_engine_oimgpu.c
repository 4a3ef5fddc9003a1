/*
 * copy_engine_oimgpu.c — liboimgpu as an SPDK copy engine.
 *
 * This is the reference-side binding for the level at which SPDK's Malloc bdev does its memcpy: a
 * `struct spdk_copy_engine {copy, fill, get_io_channel}` registered with SPDK_COPY_MODULE_REGISTER
 * (S/include/spdk_internal/copy_engine.h:47-93).  A registered "hardware" engine wins over the built-in memcpy
 * engine for every copy channel created afterwards (copy_create_cb, S/lib/copy/copy_engine.c:186-203), so an
 * UNMODIFIED bdev_malloc.c (bdev_malloc_readv/writev/unmap, S/lib/bdev/malloc/bdev_malloc.c:153-233) moves its
 * bytes through oim_copy_kernel / oim_fill_kernel once this file is linked into the SPDK application.
 *
 * Model: the I/OAT engine (S/lib/copy/ioat/copy_engine_ioat.c:114-221): submit returns at once; a poller
 * registered when the channel is created reaps completions on the channel's thread and calls
 * cb(copy_task, 0).  A submit error is returned as -errno, which bdev_malloc turns into the I/O's status
 * (malloc_done, bdev_malloc.c:73-89: -ENOMEM -> NOMEM/retry, anything else -> FAILED).
 *
 * Written against SPDK's public/internal headers as vendored by intel/oim; nothing here is copied from them.
 * Build + test: integration/spdk/Makefile, tests/test_spdk_copy_engine.py.
 */
#include "spdk/stdinc.h"

#include "spdk_internal/copy_engine.h"

#include "spdk/env.h"
#include "spdk/thread.h"
#include "spdk/log.h"
#include "spdk/string.h"

#include "oimgpu.h"

#define OIMGPU_REAP_BATCH 64

struct oimgpu_copy_task {
	spdk_copy_completion_cb	cb;
};

struct oimgpu_io_channel {
	oimgpu_copy_chan	*chan;
	struct spdk_poller	*poller;
};

static int g_oimgpu_engine_on;
static unsigned long long g_oimgpu_ops, g_oimgpu_bytes;

static int copy_engine_oimgpu_init(void);
static void copy_engine_oimgpu_exit(void *ctx);
static void copy_engine_oimgpu_config_text(FILE *fp);

static size_t
copy_engine_oimgpu_get_ctx_size(void)
{
	return sizeof(struct oimgpu_copy_task) + sizeof(struct spdk_copy_task);
}

SPDK_COPY_MODULE_REGISTER(copy_engine_oimgpu_init, copy_engine_oimgpu_exit,
			  copy_engine_oimgpu_config_text,
			  copy_engine_oimgpu_get_ctx_size)

static int
oimgpu_engine_copy(void *cb_arg, struct spdk_io_channel *ch, void *dst, void *src, uint64_t nbytes,
		   spdk_copy_completion_cb cb)
{
	struct oimgpu_copy_task *task = cb_arg;		/* = spdk_copy_task.offload_ctx of the caller's task */
	struct oimgpu_io_channel *och = spdk_io_channel_get_ctx(ch);

	task->cb = cb;
	g_oimgpu_ops++;
	g_oimgpu_bytes += nbytes;
	return oimgpu_copy_chan_copy(och->chan, dst, src, nbytes, task);
}

static int
oimgpu_engine_fill(void *cb_arg, struct spdk_io_channel *ch, void *dst, uint8_t fill, uint64_t nbytes,
		   spdk_copy_completion_cb cb)
{
	struct oimgpu_copy_task *task = cb_arg;
	struct oimgpu_io_channel *och = spdk_io_channel_get_ctx(ch);

	task->cb = cb;
	g_oimgpu_ops++;
	g_oimgpu_bytes += nbytes;
	return oimgpu_copy_chan_fill(och->chan, dst, fill, nbytes, task);
}

/* the channel's poller: completed kernels -> callbacks, on the thread that submitted them */
static int
oimgpu_engine_poll(void *arg)
{
	struct oimgpu_io_channel *och = arg;
	void *tags[OIMGPU_REAP_BATCH];
	int i, n;

	n = oimgpu_copy_chan_poll(och->chan, tags, OIMGPU_REAP_BATCH);
	for (i = 0; i < n; i++) {
		struct oimgpu_copy_task *task = tags[i];
		struct spdk_copy_task *req = (struct spdk_copy_task *)((uintptr_t)task - offsetof(struct spdk_copy_task, offload_ctx));

		task->cb(req, 0);
	}
	return n > 0 ? n : -1;
}

static struct spdk_io_channel *oimgpu_engine_get_io_channel(void);

static struct spdk_copy_engine oimgpu_copy_engine = {
	.copy		= oimgpu_engine_copy,
	.fill		= oimgpu_engine_fill,
	.get_io_channel	= oimgpu_engine_get_io_channel,
};

static int
oimgpu_engine_create_cb(void *io_device, void *ctx_buf)
{
	struct oimgpu_io_channel *och = ctx_buf;

	if (oimgpu_copy_chan_open(-1, &och->chan) != 0) {
		return -1;	/* copy_create_cb then falls back to the memcpy engine for this channel */
	}
	och->poller = spdk_poller_register(oimgpu_engine_poll, och, 0);
	return 0;
}

static void
oimgpu_engine_destroy_cb(void *io_device, void *ctx_buf)
{
	struct oimgpu_io_channel *och = ctx_buf;

	spdk_poller_unregister(&och->poller);
	oimgpu_copy_chan_close(och->chan);
}

static struct spdk_io_channel *
oimgpu_engine_get_io_channel(void)
{
	return spdk_get_io_channel(&oimgpu_copy_engine);
}

static int
copy_engine_oimgpu_init(void)
{
	const char *off = getenv("OIMGPU_COPY_ENGINE");
	int rc;

	if (off && strcmp(off, "off") == 0) {
		return 0;	/* "Users may not want to use offload even it is available" (the I/OAT engine's Enable No) */
	}
	rc = oimgpu_init(NULL, 0);
	if (rc != 0) {
		SPDK_NOTICELOG("oimgpu copy engine: no usable GPU (%s), memcpy engine stays\n", spdk_strerror(-rc));
		return 0;
	}
	spdk_copy_engine_register(&oimgpu_copy_engine);
	spdk_io_device_register(&oimgpu_copy_engine, oimgpu_engine_create_cb, oimgpu_engine_destroy_cb,
				sizeof(struct oimgpu_io_channel), "oimgpu_copy_engine");
	g_oimgpu_engine_on = 1;
	return 0;
}

static void
copy_engine_oimgpu_exit(void *ctx)
{
	if (g_oimgpu_engine_on) {
		spdk_io_device_unregister(&oimgpu_copy_engine, NULL);
		g_oimgpu_engine_on = 0;
	}
	spdk_copy_engine_module_finish();
}

static void
copy_engine_oimgpu_config_text(FILE *fp)
{
	fprintf(fp, "[OimGpu]\n  Enable %s\n", g_oimgpu_engine_on ? "Yes" : "No");
}

/* for the test driver: is the engine the one SPDK picked, and how much went through it */
int copy_engine_oimgpu_active(void) { return g_oimgpu_engine_on; }
void copy_engine_oimgpu_counters(unsigned long long *ops, unsigned long long *bytes) { *ops = g_oimgpu_ops; *bytes = g_oimgpu_bytes; }
