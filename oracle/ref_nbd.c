/*
 * ref_nbd.c — TEST INFRASTRUCTURE.  Drives the REFERENCE's NBD server loop (S/lib/nbd/nbd.c: request
 * framing, spdk_bdev_read/write/flush/unmap, reply framing) without the kernel half: spdk_nbd_start()
 * hands one end of a socketpair to /dev/nbdX and polls the other; here the caller keeps the "kernel"
 * end itself and speaks the kernel's side of the protocol (struct nbd_request / nbd_reply,
 * linux/nbd.h), so nothing needs /dev/nbd* or CAP_SYS_ADMIN.  nbd.c is #included because the disk
 * structure and the poll function are private to it (SPDK's own unit tests include the .c under test).
 */
#include "nbd/nbd.c"

#include <fcntl.h>

void oimref_enter(void);	/* ref_driver.c: make the calling thread the SPDK thread */

void spdk_unaffinitize_thread(void) {}	/* env call of the kernel-attach thread, never started here */

static void ref_nbd_hot_remove(void *ctx) { spdk_nbd_stop(ctx); }

/* The state spdk_nbd_start() + spdk_nbd_start_complete() leave behind (nbd.c:893-1060), minus the ioctls:
 * bdev opened, channel taken, socket non-blocking, poller registered.  fd is dup()ed. */
void *oimrefnbd_start(const char *bdev_name, int fd)
{
	struct spdk_nbd_disk *nbd;
	struct spdk_bdev *bdev;
	int flag;

	static int inited;

	oimref_enter();
	if (!inited) { spdk_nbd_init(); inited = 1; }
	bdev = spdk_bdev_get_by_name(bdev_name);
	if (!bdev) return NULL;
	nbd = calloc(1, sizeof(*nbd));
	nbd->dev_fd = -1;
	nbd->kernel_sp_fd = -1;
	nbd->spdk_sp_fd = dup(fd);
	if (spdk_bdev_open(bdev, true, ref_nbd_hot_remove, nbd, &nbd->bdev_desc) != 0) { free(nbd); return NULL; }
	nbd->bdev = bdev;
	nbd->ch = spdk_bdev_get_io_channel(nbd->bdev_desc);
	nbd->buf_align = spdk_max(spdk_bdev_get_buf_align(bdev), 64);
	nbd->nbd_path = strdup("/dev/nbd-none");
	TAILQ_INIT(&nbd->received_io_list);
	TAILQ_INIT(&nbd->executed_io_list);
	if (spdk_nbd_disk_register(nbd) != 0) return NULL;
	flag = fcntl(nbd->spdk_sp_fd, F_GETFL);
	fcntl(nbd->spdk_sp_fd, F_SETFL, flag | O_NONBLOCK);
	nbd->nbd_poller = spdk_poller_register(spdk_nbd_poll, nbd, 0);
	return nbd;
}

/* 1 while the disk is registered (the poller closes and frees it on disconnect / protocol error) */
int oimrefnbd_alive(void)
{
	return spdk_nbd_disk_find_by_nbd_path("/dev/nbd-none") != NULL;
}
