/*
 * vhost_user.cpp — vhost-user slave for vhost-scsi controllers served by the GPU.  See vhost_user.h.
 *
 * Protocol behaviour follows S/lib/vhost/rte_vhost/vhost_user.c (message by message; each handler
 * below names the function it mirrors), device start/stop follows S/lib/vhost/vhost.c:994-1134, the
 * control / event queues follow S/lib/vhost/vhost_scsi.c:236-482.  tests/test_vhost_user.py replays one
 * master script against this slave and against the reference's own (oracle/_ref/liboim_ref_vhost.so) and
 * compares the transcripts and the guest memory.
 */
#include "vhost_user.h"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include <execinfo.h>
#include <fcntl.h>
#include <poll.h>
#include <pthread.h>
#include <signal.h>
#include <sys/eventfd.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

#include "oimgpu.h"

namespace vhost_user {
namespace {

/* VhostUserRequest (vhost_user.h:65-96) */
enum : uint32_t {
	GET_FEATURES = 1, SET_FEATURES = 2, SET_OWNER = 3, RESET_OWNER = 4, SET_MEM_TABLE = 5, SET_LOG_BASE = 6,
	SET_LOG_FD = 7, SET_VRING_NUM = 8, SET_VRING_ADDR = 9, SET_VRING_BASE = 10, GET_VRING_BASE = 11,
	SET_VRING_KICK = 12, SET_VRING_CALL = 13, SET_VRING_ERR = 14, GET_PROTOCOL_FEATURES = 15,
	SET_PROTOCOL_FEATURES = 16, GET_QUEUE_NUM = 17, SET_VRING_ENABLE = 18, SEND_RARP = 19, NET_SET_MTU = 20,
	GET_CONFIG = 24, SET_CONFIG = 25, REQ_MAX = 86,
};
constexpr uint32_t kVersion = 0x1, kVersionMask = 0x3, kReplyMask = 0x4, kNeedReply = 0x8;
constexpr uint64_t kVringIdxMask = 0xff, kVringNoFd = 0x100;
constexpr uint32_t kMaxVring = 0x100, kMaxQueuePairs = 0x80, kMaxRegions = 8;
constexpr uint32_t kMaxPayload = 64 + 4096;	/* the largest member of VhostUserMsg.payload (nvme) */
/* SPDK_VHOST_SCSI_FEATURES & ~SPDK_VHOST_SCSI_DISABLED_FEATURES (vhost_scsi.c:51-62, vhost_internal.h:86-94):
 * VIRTIO_SCSI_F_INOUT|HOTPLUG|CHANGE, VHOST_F_LOG_ALL(26), INDIRECT_DESC(28), PROTOCOL_FEATURES(30), VERSION_1(32) */
/* What the reference offers (0x154000007 / 0x21f) minus dirty-page logging: VHOST_F_LOG_ALL (26) and the LOG_SHMFD
 * protocol feature (1) are NOT offered.  The device writes guest memory from the GPU; logging those writes into the
 * master's bitmap (vhost_user_set_log_base, rte_vhost/vhost_user.c:899-960; vhost_log_write, vhost.h) would need
 * system-scope atomics into host memory that must not lose a bit against QEMU's concurrent test-and-clear, which PCIe
 * only guarantees with AtomicOps routing.  Offering the bit without logging would let a live migration lose writes
 * silently; without it QEMU registers a migration blocker, which is the honest answer.  A master that sets the
 * bits anyway is accepted as the reference accepts it (nothing is logged). */
constexpr uint64_t kFeaturesRef = 0x7ull | 1ull << 26 | 1ull << 28 | 1ull << 30 | 1ull << 32;
constexpr uint64_t kFeatures = kFeaturesRef & ~(1ull << 26);
constexpr uint64_t kProtocolFeaturesRef = 0x21f;	/* MQ, LOG_SHMFD, RARP, REPLY_ACK, NET_MTU, CONFIG (vhost_user.h:51-63) */
constexpr uint64_t kProtocolFeatures = kProtocolFeaturesRef & ~(1ull << 1);
constexpr int kFdUninit = -1, kFdInvalid = -2;	/* VIRTIO_UNINITIALIZED_EVENTFD / VIRTIO_INVALID_EVENTFD */
constexpr uint64_t kMask2M = (2ull << 20) - 1;

enum : uint16_t { VD_NEXT = 1, VD_WRITE = 2, VD_INDIRECT = 4 };
struct VringDesc { uint64_t addr; uint32_t len; uint16_t flags, next; };
struct UsedElem { uint32_t id, len; };

struct Vq {
	uint32_t size = 0;
	uint16_t last_avail = 0, last_used = 0;
	uint8_t *desc = nullptr, *avail = nullptr, *used = nullptr;	/* host addresses inside a region mapping */
	int kickfd = kFdUninit, callfd = kFdUninit;
	bool enabled = false;
	bool attached = false;		/* handed to the GPU for the current run */
	uint16_t consumed = 0;		/* avail idx the GPU was last kicked for */
	uint16_t signalled = 0;		/* used idx the guest was last interrupted for */
	/* interrupt coalescing (S/lib/vhost/vhost.c:270-340): spdk_vhost_virtqueue::req_cnt / irq_delay_time / next_event_time */
	uint16_t stat_used = 0;		/* used idx at the last statistics check */
	uint32_t req_cnt = 0;
	uint64_t irq_delay_ns = 0, next_event_ns = 0;
};

struct Region {
	uint64_t gpa = 0, size = 0, uva = 0, mmap_off = 0;
	int fd = -1;
	void *mmap_addr = nullptr;
	uint64_t mmap_size = 0;
	uint8_t *host = nullptr;	/* mmap_addr + mmap_off */
	uint64_t dev = 0;		/* the same bytes as the GPU addresses them */
	bool registered = false;
};

struct MemTableMsg {
	uint32_t nregions, padding;
	struct { uint64_t gpa, size, uva, mmap_off; } regions[kMaxRegions];
};

Config g_cfg;
std::mutex g_mu;
const bool g_debug = getenv("OIM_VU_DEBUG") != nullptr;
#define VU_DEBUG(...) do { if (g_debug) { fprintf(stderr, "vhost-user: " __VA_ARGS__); fputc('\n', stderr); } } while (0)

/* diagnostics: `kill -USR2 <pid>` makes every transport thread print where it is (stderr) */
std::mutex g_threads_mu;
std::vector<pthread_t> g_threads;

void on_usr1(int)
{
	void *bt[40];
	const int n = backtrace(bt, 40);
	backtrace_symbols_fd(bt, n, 2);
	const char nl[] = "----\n";
	(void)!write(2, nl, sizeof(nl) - 1);
}

void on_usr2(int)
{
	std::lock_guard<std::mutex> lk(g_threads_mu);
	for (pthread_t t : g_threads) pthread_kill(t, SIGUSR1);
}

void register_thread()
{
	std::lock_guard<std::mutex> lk(g_threads_mu);
	g_threads.push_back(pthread_self());
}

void unregister_thread()
{
	std::lock_guard<std::mutex> lk(g_threads_mu);
	for (size_t i = 0; i < g_threads.size(); i++) {
		if (pthread_equal(g_threads[i], pthread_self())) { g_threads.erase(g_threads.begin() + i); break; }
	}
}

struct Server;

struct Session {
	Server *srv = nullptr;
	int fd = -1;
	int wake = -1;
	std::thread th;
	std::atomic<bool> done{false};

	uint64_t features = 0, protocol_features = 0;
	bool has_new_table = false;
	MemTableMsg new_table{};
	int new_fds[kMaxRegions];
	std::vector<Region> mem;
	bool have_mem = false;
	std::vector<Vq> vq;
	bool running = false;
	uint32_t max_queues = 0;
	/* One data-path session per GPU the daemon owns: request queue r is served by luns[r % G] as its queue r / G.
	 * OIM attaches every volume as another target of the ONE controller (pkg/oim-controller/controller.go:131-148) and
	 * the targets may live on different GPUs; any GPU reaches any target (its own HBM or a peer's over NVLink), so
	 * dealing out the QUEUES spreads the guest-memory traffic over every GPU's PCIe link and keeps each used ring
	 * single-writer. */
	std::vector<oimgpu_lun *> luns;
	oimgpu_lun *lun = nullptr;	/* luns[0]: "the data path is open" */
	uint32_t lun_queues = 0;
	void close_luns();
	bool polling = false;
	std::mutex ev_mu;
	std::vector<std::pair<int, bool>> events;	/* hot-plug notifications from the RPC thread */
	uint64_t next_stats_check_ns = 0;		/* check_session_io_stats: every SPDK_VHOST_STATS_CHECK_INTERVAL_MS = 10 ms */
	uint32_t coalescing_delay_us = 0, coalescing_iops = 60000;
	unsigned long long irqs_sent = 0, irqs_held = 0;
	void check_io_stats(uint64_t now);

	void run();
	bool handle_message();
	void teardown();
	bool is_ready() const;
	bool start();
	void stop();
	void service();
	void free_mem();
	int setup_mem();
	uint8_t *qva_to_host(uint64_t qva, uint64_t *len) const;
	uint8_t *gpa_to_host(uint64_t gpa, uint64_t len) const;
	uint64_t host_to_dev(const uint8_t *p) const;
	void used_enqueue(Vq &q, uint16_t id, uint32_t len);
	int avail_get(Vq &q, uint16_t *reqs, int max);
	bool get_desc(Vq &q, uint16_t head, VringDesc **desc, VringDesc **table, uint32_t *table_size);
	void process_controlq();
	void eventq_enqueue(int target, uint32_t event, uint32_t reason);
	bool target_present(const uint8_t *lun) const;
	void signal_used(Vq &q);
	bool reply(uint32_t req, uint32_t flags, const void *payload, uint32_t size);
};

struct Server {
	std::string name, path;
	int lfd = -1;
	int wake = -1;
	std::thread th;
	std::atomic<bool> stop{false};
	std::mutex mu;
	std::vector<std::unique_ptr<Session>> sessions;
	void accept_loop();
};

std::map<std::string, std::unique_ptr<Server>> g_servers;

/* ---- wire ------------------------------------------------------------------------------------ */

/* read_vhost_message (vhost_user.c:1002-1029): header + fds in one recvmsg, then the payload.
 * -> 1 ok, 0 peer closed, -1 error */
int read_message(int fd, uint32_t hdr[3], std::vector<uint8_t> &payload, int fds[kMaxRegions], int *nfds)
{
	struct iovec iov = {hdr, 12};
	char control[CMSG_SPACE(kMaxRegions * sizeof(int))];
	struct msghdr mh {};
	mh.msg_iov = &iov;
	mh.msg_iovlen = 1;
	mh.msg_control = control;
	mh.msg_controllen = sizeof(control);
	*nfds = 0;
	ssize_t n;
	do { n = recvmsg(fd, &mh, MSG_CMSG_CLOEXEC); } while (n < 0 && errno == EINTR);
	if (n == 0) return 0;
	if (n != 12 || (mh.msg_flags & (MSG_TRUNC | MSG_CTRUNC))) return -1;
	for (struct cmsghdr *c = CMSG_FIRSTHDR(&mh); c; c = CMSG_NXTHDR(&mh, c)) {
		if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS) {
			int k = (int)((c->cmsg_len - CMSG_LEN(0)) / sizeof(int));
			if (k > (int)kMaxRegions) k = kMaxRegions;
			memcpy(fds, CMSG_DATA(c), k * sizeof(int));
			*nfds = k;
		}
	}
	if (hdr[2] > kMaxPayload) return -1;
	payload.assign(hdr[2], 0);
	size_t got = 0;
	while (got < hdr[2]) {
		ssize_t r = read(fd, payload.data() + got, hdr[2] - got);
		if (r < 0 && errno == EINTR) continue;
		if (r <= 0) return -1;
		got += (size_t)r;
	}
	return 1;
}

/* send_vhost_message (vhost_user.c:1031-1048) */
bool Session::reply(uint32_t req, uint32_t flags, const void *payload, uint32_t size)
{
	std::vector<uint8_t> out(12 + size);
	uint32_t h[3] = {req, (flags & ~kVersionMask & ~kNeedReply) | kVersion | kReplyMask, size};
	memcpy(out.data(), h, 12);
	if (size) memcpy(out.data() + 12, payload, size);
	size_t sent = 0;
	while (sent < out.size()) {
		ssize_t n = send(fd, out.data() + sent, out.size() - sent, MSG_NOSIGNAL);
		if (n < 0 && errno == EINTR) continue;
		if (n <= 0) return false;
		sent += (size_t)n;
	}
	return true;
}

/* ---- guest memory ------------------------------------------------------------------------------ */

void Session::free_mem()
{
	for (Region &r : mem) {
		if (r.registered) oimgpu_mem_unregister(r.mmap_addr);
		if (r.mmap_addr) munmap(r.mmap_addr, r.mmap_size);
		if (r.fd >= 0) close(r.fd);
	}
	mem.clear();
	have_mem = false;
}

/* vhost_setup_mem_table (vhost_user.c:587-712): map every region of the pending table; the ring
 * addresses of all queues become invalid and must be sent again */
int Session::setup_mem()
{
	free_mem();
	for (Vq &q : vq) q.desc = q.avail = q.used = nullptr;
	for (uint32_t i = 0; i < new_table.nregions && i < kMaxRegions; i++) {
		Region r;
		r.gpa = new_table.regions[i].gpa;
		r.size = new_table.regions[i].size;
		r.uva = new_table.regions[i].uva;
		r.mmap_off = new_table.regions[i].mmap_off;
		r.fd = new_fds[i];
		new_fds[i] = -1;
		struct stat st;
		if (fstat(r.fd, &st) != 0) { mem.push_back(r); free_mem(); return -1; }
		const uint64_t align = (uint64_t)st.st_blksize;	/* the hugepage size on hugetlbfs */
		r.mmap_size = (r.size + r.mmap_off + align - 1) / align * align;
		void *p = mmap(nullptr, r.mmap_size, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, r.fd, 0);
		if (p == MAP_FAILED) { mem.push_back(r); free_mem(); return -1; }
		madvise(p, r.mmap_size, MADV_DONTDUMP);
		r.mmap_addr = p;
		r.host = (uint8_t *)p + r.mmap_off;
		r.dev = (uint64_t)(uintptr_t)r.host;
		mem.push_back(r);
	}
	for (uint32_t i = new_table.nregions; i < kMaxRegions; i++) {
		if (new_fds[i] >= 0) { close(new_fds[i]); new_fds[i] = -1; }
	}
	have_mem = true;
	return 0;
}

/* qva_to_vva (vhost_user.c:337-359): master virtual address -> ours; *len is clipped to the region */
uint8_t *Session::qva_to_host(uint64_t qva, uint64_t *len) const
{
	for (const Region &r : mem) {
		if (qva >= r.uva && qva < r.uva + r.size) {
			if (*len > r.uva + r.size - qva) *len = r.uva + r.size - qva;
			return r.host + (qva - r.uva);
		}
	}
	return nullptr;
}

/* spdk_vhost_gpa_to_vva (vhost.c:93-106) */
uint8_t *Session::gpa_to_host(uint64_t gpa, uint64_t len) const
{
	for (const Region &r : mem) {
		if (gpa >= r.gpa && gpa < r.gpa + r.size) {
			if (len > r.gpa + r.size - gpa) return nullptr;
			return r.host + (gpa - r.gpa);
		}
	}
	return nullptr;
}

uint64_t Session::host_to_dev(const uint8_t *p) const
{
	for (const Region &r : mem) {
		if (p >= r.host && p < r.host + r.size) return r.dev + (uint64_t)(p - r.host);
	}
	return 0;
}

/* ---- split-ring helpers for the two CPU-served queues -------------------------------------------- */

/* spdk_vhost_vq_avail_ring_get (vhost.c:178-211) */
int Session::avail_get(Vq &q, uint16_t *reqs, int max)
{
	const uint16_t avail_idx = *(volatile uint16_t *)(q.avail + 2);
	uint16_t count = (uint16_t)(avail_idx - q.last_avail);
	if (count == 0) return 0;
	if (count > q.size) return 0;	/* "the queue is unrecoverably broken" */
	if (count > max) count = (uint16_t)max;
	std::atomic_thread_fence(std::memory_order_acquire);
	const uint16_t *ring = (const uint16_t *)(q.avail + 4);
	for (uint16_t i = 0; i < count; i++) reqs[i] = ring[(q.last_avail + i) & (q.size - 1)];
	q.last_avail = (uint16_t)(q.last_avail + count);
	return count;
}

/* spdk_vhost_vq_get_desc (vhost.c:219-247) */
bool Session::get_desc(Vq &q, uint16_t head, VringDesc **desc, VringDesc **table, uint32_t *table_size)
{
	if (head >= q.size) return false;
	*desc = (VringDesc *)q.desc + head;
	if ((*desc)->flags & VD_INDIRECT) {
		*table_size = (*desc)->len / sizeof(VringDesc);
		if (*table_size == 0) return false;	/* the reference would read a descriptor out of a 0-byte table */
		*table = (VringDesc *)gpa_to_host((*desc)->addr, sizeof(VringDesc) * (uint64_t)*table_size);
		*desc = *table;
		return *desc != nullptr;
	}
	*table = (VringDesc *)q.desc;
	*table_size = q.size;
	return true;
}

/* spdk_vhost_vring_desc_get_next (vhost.c:433-453); false + *desc == nullptr: end or bad index */
static bool desc_next(VringDesc **desc, VringDesc *table, uint32_t table_size)
{
	VringDesc *d = *desc;
	if (!(d->flags & VD_NEXT)) { *desc = nullptr; return true; }
	if (d->next >= table_size) { *desc = nullptr; return false; }
	*desc = &table[d->next];
	return true;
}

/* spdk_vhost_vq_used_ring_enqueue (vhost.c:397-431) */
void Session::used_enqueue(Vq &q, uint16_t id, uint32_t len)
{
	UsedElem *ring = (UsedElem *)(q.used + 4);
	const uint16_t slot = q.last_used & (q.size - 1);
	q.last_used++;
	ring[slot].id = id;
	ring[slot].len = len;
	std::atomic_thread_fence(std::memory_order_release);
	*(volatile uint16_t *)(q.used + 2) = q.last_used;
	std::atomic_thread_fence(std::memory_order_seq_cst);
}

/* spdk_vhost_vq_used_signal (vhost.c:249-266): one interrupt for everything completed since the last one */
static uint64_t now_ns()
{
	timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

/* check_session_io_stats (vhost.c:270-297): every 10 ms, a queue whose request count since it last crossed the
 * threshold exceeds the threshold gets an interrupt delay proportional to the excess; the count is only reset
 * when it does (the reference's arithmetic, kept as is) */
void Session::check_io_stats(uint64_t now)
{
	if (now < next_stats_check_ns) return;
	next_stats_check_ns = now + 10000000ull;
	oimgpu_ctrlr_info info;
	if (oimgpu_vhost_ctrlr_get(srv->name.c_str(), &info) == 0) {	/* set_vhost_controller_coalescing reaches running sessions */
		coalescing_delay_us = info.delay_base_us;
		coalescing_iops = info.iops_threshold;
	}
	const uint32_t io_threshold = coalescing_iops * 10u / 1000u;
	if (!coalescing_delay_us || !io_threshold) return;
	for (Vq &q : vq) {
		if (!q.used) continue;
		const uint16_t idx = *(volatile uint16_t *)(q.used + 2);
		q.req_cnt += (uint16_t)(idx - q.stat_used);
		q.stat_used = idx;
		if (q.req_cnt <= io_threshold) continue;
		q.irq_delay_ns = (uint64_t)coalescing_delay_us * 1000ull * (q.req_cnt - io_threshold) / io_threshold;
		q.req_cnt = 0;
		q.next_event_ns = now;
	}
}

/* spdk_vhost_session_used_signal / spdk_vhost_vq_used_signal (vhost.c:249-340): an interrupt when the used index
 * moved, unless the guest masked them (VRING_AVAIL_F_NO_INTERRUPT) or the queue's coalescing delay has not passed */
void Session::signal_used(Vq &q)
{
	if (!q.used) return;
	const uint16_t idx = *(volatile uint16_t *)(q.used + 2);
	if (idx == q.signalled) return;
	if (q.avail && (*(volatile uint16_t *)q.avail & 1)) return;	/* stays pending: signalled is not advanced */
	if (coalescing_delay_us) {
		const uint64_t now = now_ns();
		if (now < q.next_event_ns) { irqs_held++; return; }
		q.next_event_ns = now + q.irq_delay_ns;
	}
	q.signalled = idx;
	irqs_sent++;
	if (q.callfd >= 0) eventfd_write(q.callfd, 1);
}

/* spdk_vhost_scsi_task_init_target (vhost_scsi.c:361-387), presence only */
bool Session::target_present(const uint8_t *lun) const
{
	if (lun[0] != 1 || lun[1] >= OIMGPU_CTRLR_MAX_DEVS) return false;
	oimgpu_ctrlr_info info;
	if (oimgpu_vhost_ctrlr_get(srv->name.c_str(), &info) != 0) return false;
	for (uint32_t i = 0; i < info.ntargets; i++) {
		if (info.targets[i].scsi_dev_num == lun[1]) return true;
	}
	return false;
}

/* process_controlq + process_ctrl_request (vhost_scsi.c:389-482, 655-687): task management and
 * asynchronous-notification requests, a few per VM lifetime - served on the CPU */
void Session::process_controlq()
{
	if (vq.size() < 1 || !vq[0].desc || !vq[0].size) return;
	Vq &q = vq[0];
	uint16_t reqs[32];
	const int n = avail_get(q, reqs, 32);
	for (int i = 0; i < n; i++) {
		const uint16_t head = reqs[i];
		uint32_t used_len = 0;
		VringDesc *desc = nullptr, *table = nullptr;
		uint32_t table_size = 0;
		do {
			if (head >= q.size) break;
			if (!get_desc(q, head, &desc, &table, &table_size)) break;
			/* struct virtio_scsi_ctrl_tmf_req { u32 type; u32 subtype; u8 lun[8]; u64 tag; } */
			const uint8_t *req = gpa_to_host(desc->addr, 24);
			if (!req) break;
			uint32_t type, subtype;
			memcpy(&type, req, 4);
			memcpy(&subtype, req + 4, 4);
			const bool present = target_present(req + 8);
			desc_next(&desc, table, table_size);
			if (!desc) break;	/* no response descriptor */
			if (type == 0) {		/* VIRTIO_SCSI_T_TMF */
				uint8_t *resp = gpa_to_host(desc->addr, 1);
				if (desc->len < 1 || !resp) break;
				if (!present) {
					*resp = OIMGPU_S_BAD_TARGET;
				} else if (subtype == 5) {	/* VIRTIO_SCSI_T_TMF_LOGICAL_UNIT_RESET */
					/* mgmt_task_submit (vhost_scsi.c:340-346): everything in flight on the device
					 * finishes, then FUNCTION COMPLETE; the used length of a management task is
					 * whatever the slot's task last carried - zero */
					if (lun && !polling) { for (oimgpu_lun *l : luns) oimgpu_lun_sync(l); }
					*resp = 0;	/* VIRTIO_SCSI_S_OK */
					break;		/* used_len stays 0 */
				} else {
					*resp = 2;	/* VIRTIO_SCSI_S_ABORTED */
				}
			} else if (type == 1 || type == 2) {	/* VIRTIO_SCSI_T_AN_QUERY / _SUBSCRIBE */
				/* struct virtio_scsi_ctrl_an_resp { u32 event_actual; u8 response; } */
				uint8_t *resp = gpa_to_host(desc->addr, 5);
				if (desc->len < 5 || !resp) break;
				resp[4] = 2;
			}
			used_len = 1;	/* sizeof(struct virtio_scsi_ctrl_tmf_resp) */
		} while (0);
		used_enqueue(q, head, used_len);
	}
	if (n) signal_used(q);
}

/* eventq_enqueue (vhost_scsi.c:236-289) */
void Session::eventq_enqueue(int target, uint32_t event, uint32_t reason)
{
	if (vq.size() < 2 || !vq[1].desc || !vq[1].size) return;
	Vq &q = vq[1];
	uint16_t head;
	if (avail_get(q, &head, 1) != 1) {
		fprintf(stderr, "oim-gpu-vhost: controller %s: no event-queue buffer for a hot-plug event\n", srv->name.c_str());
		return;
	}
	uint32_t size = 0;
	VringDesc *desc = nullptr, *table = nullptr;
	uint32_t table_size = 0;
	if (get_desc(q, head, &desc, &table, &table_size) && desc->len >= 16) {
		/* struct virtio_scsi_event { u32 event; u8 lun[8]; u32 reason; } */
		uint8_t *ev = gpa_to_host(desc->addr, 16);
		if (ev) {
			memcpy(ev, &event, 4);
			memset(ev + 4, 0, 8);
			ev[4] = 1;
			ev[5] = (uint8_t)target;
			memcpy(ev + 12, &reason, 4);
			size = 16;
		}
	}
	used_enqueue(q, head, size);
	signal_used(q);
}

/* ---- device start / stop ------------------------------------------------------------------------- */

/* vq_is_ready / virtio_is_ready (vhost_user.c:714-744): ONE usable ring is enough */
bool Session::is_ready() const
{
	for (const Vq &q : vq) {
		if (q.desc && q.kickfd != kFdUninit && q.callfd != kFdUninit && q.kickfd != kFdInvalid && q.callfd != kFdInvalid)
			return true;
	}
	return false;
}

/* start_device (vhost.c:1044-1134) + spdk_vhost_scsi_start (vhost_scsi.c:1236-1292) */
bool Session::start()
{
	max_queues = 0;
	for (uint32_t i = 0; i < vq.size(); i++) {
		if (!vq[i].desc || !vq[i].size) continue;
		/* the reference always tells the guest not to kick (it polls); only the resident-poller mode
		 * polls here, a launch per kick needs the kick */
		uint16_t fl = g_cfg.poller ? 1 : 0;	/* VRING_USED_F_NO_NOTIFY */
		memcpy(vq[i].used, &fl, 2);
		max_queues = i + 1;
	}
	for (const Region &r : mem) {
		if (r.size & kMask2M) {
			fprintf(stderr, "oim-gpu-vhost: %s: guest memory size is not a 2MB multiple\n", srv->name.c_str());
			return false;
		}
		/* beyond the reference: its 2 MiB-page walk (vhost.c:461-509) stays inside a region only if the region
		 * also STARTS on a 2 MiB boundary, which QEMU guarantees for hugepage-backed memory; on a GPU a stray
		 * access is not a SIGSEGV of one request but the end of the context, so insist on it */
		if (r.gpa & kMask2M) {
			fprintf(stderr, "oim-gpu-vhost: %s: guest memory region does not start on a 2MB boundary\n", srv->name.c_str());
			return false;
		}
	}
	for (uint32_t i = 0; i < max_queues; i++) {
		if (vq[i].callfd >= 0) eventfd_write(vq[i].callfd, 1);	/* vhost.c:1101-1115 */
		vq[i].signalled = vq[i].used ? *(volatile uint16_t *)(vq[i].used + 2) : 0;
		vq[i].consumed = vq[i].last_avail;
	}
	if (!g_cfg.control_only) {
		for (Region &r : mem) {
			if (r.registered) continue;
			int rc = oimgpu_mem_register(r.mmap_addr, r.mmap_size);
			if (rc != 0 && rc != -EEXIST) {
				fprintf(stderr, "oim-gpu-vhost: %s: cannot pin guest memory: %s\n", srv->name.c_str(), strerror(-rc));
				return false;
			}
			r.registered = rc == 0;
			uint64_t dev = 0;
			if (oimgpu_mem_device_addr(r.host, &dev) != 0) {
				fprintf(stderr, "oim-gpu-vhost: %s: guest memory has no device address\n", srv->name.c_str());
				return false;
			}
			r.dev = dev;
		}
		const uint32_t want = max_queues > 2 ? max_queues - 2 : 1;
		const uint32_t ngpu = g_cfg.spread ? std::max(1, std::min(oimgpu_device_count(), (int)want)) : 1;
		if (lun && (lun_queues != want || luns.size() != ngpu)) close_luns();
		if (!lun) {
			for (uint32_t gidx = 0; gidx < ngpu; gidx++) {
				oimgpu_lun *l = nullptr;
				const uint32_t nq = (want + ngpu - 1 - gidx) / ngpu;	/* queues gidx, gidx + G, ... */
				int rc = ngpu == 1 ? oimgpu_lun_open(srv->name.c_str(), -1, want, 32, &l)
						   : oimgpu_lun_open_on(srv->name.c_str(), oimgpu_device_ordinal((int)gidx), nq ? nq : 1, 32, &l);
				if (rc != 0) {
					fprintf(stderr, "oim-gpu-vhost: %s: cannot open the data path: %s\n", srv->name.c_str(), strerror(-rc));
					close_luns();
					return false;
				}
				luns.push_back(l);
			}
			lun = luns[0];
			lun_queues = want;
		}
		VU_DEBUG("%s: session %d: data path open on %zu GPU(s)", srv->name.c_str(), fd, luns.size());
		std::vector<oimgpu_mem_region> tbl;
		for (const Region &r : mem) tbl.push_back({r.gpa, r.size, r.dev});
		for (oimgpu_lun *l : luns) {
			int mrc = oimgpu_lun_set_mem_table(l, tbl.data(), (uint32_t)tbl.size());
			if (mrc != 0) {
				fprintf(stderr, "oim-gpu-vhost: %s: memory table refused: %s\n", srv->name.c_str(), strerror(-mrc));
				return false;
			}
		}
		for (uint32_t i = 2; i < max_queues; i++) {
			Vq &q = vq[i];
			q.attached = false;
			if (!q.desc || !q.size) continue;
			const uint32_t r = i - 2, G = (uint32_t)luns.size();
			int rc = oimgpu_vq_attach(luns[r % G], r / G, (void *)(uintptr_t)host_to_dev(q.desc), (void *)(uintptr_t)host_to_dev(q.avail),
						  (void *)(uintptr_t)host_to_dev(q.used), q.size, q.last_avail, q.last_used);
			if (rc != 0) {
				fprintf(stderr, "oim-gpu-vhost: %s: queue %u: %s\n", srv->name.c_str(), i, strerror(-rc));
				continue;
			}
			q.attached = true;
		}
		VU_DEBUG("%s: session %d: rings attached", srv->name.c_str(), fd);
		if (g_cfg.poller) {
			polling = true;
			for (oimgpu_lun *l : luns) {
				int rc = oimgpu_lun_start_poller(l, 0, 0);
				if (rc < 0) {
					fprintf(stderr, "oim-gpu-vhost: %s: resident poller: %s\n", srv->name.c_str(), strerror(-rc));
					polling = false;
				}
			}
			if (!polling) { for (oimgpu_lun *l : luns) oimgpu_lun_stop_poller(l); }
		}
	}
	running = true;
	VU_DEBUG("%s: session %d started: %u queues, %zu regions, %s", srv->name.c_str(), fd, max_queues, mem.size(),
		 polling ? "resident poller" : lun ? "launch per kick" : "no data path");
	return true;
}

void Session::close_luns()
{
	for (oimgpu_lun *l : luns) oimgpu_lun_close(l);
	luns.clear();
	lun = nullptr;
}

/* stop_device (vhost.c:994-1041): quiesce, remember where every ring stopped */
void Session::stop()
{
	if (!running) return;
	VU_DEBUG("%s: session %d: stopping", srv->name.c_str(), fd);
	if (lun) {
		if (polling) { for (oimgpu_lun *l : luns) oimgpu_lun_stop_poller(l); polling = false; }
		for (oimgpu_lun *l : luns) oimgpu_lun_sync(l);
		for (uint32_t i = 2; i < vq.size(); i++) {
			if (!vq[i].attached) continue;
			uint16_t la = 0, lu = 0;
			const uint32_t r = i - 2, G = (uint32_t)luns.size();
			if (oimgpu_vq_detach(luns[r % G], r / G, &la, &lu) == 0) { vq[i].last_avail = la; vq[i].last_used = lu; }
			vq[i].attached = false;
		}
	}
	running = false;
	VU_DEBUG("%s: session %d stopped", srv->name.c_str(), fd);
}

/* vdev_worker + vdev_mgmt_worker (vhost_scsi.c:742-772), event driven */
void Session::service()
{
	if (!running) return;
	{
		std::vector<std::pair<int, bool>> evs;
		{
			std::lock_guard<std::mutex> lk(ev_mu);
			evs.swap(events);
		}
		for (auto &e : evs) {
			if (features & (1ull << 1))	/* VIRTIO_SCSI_F_HOTPLUG */
				eventq_enqueue(e.first, 1 /* T_TRANSPORT_RESET */, e.second ? 1 /* RESCAN */ : 2 /* REMOVED */);
		}
	}
	process_controlq();
	if (!lun) return;
	check_io_stats(now_ns());
	if (!polling) {
		for (int round = 0; round < 64; round++) {
			bool work = false;
			for (uint32_t i = 2; i < vq.size(); i++) {
				Vq &q = vq[i];
				if (!q.attached) continue;
				const uint16_t a = *(volatile uint16_t *)(q.avail + 2);
				if (a != q.consumed) { q.consumed = a; work = true; }
			}
			if (!work) break;
			bool failed = false;
			for (oimgpu_lun *l : luns) failed |= oimgpu_vq_kick(l) < 0;	/* every GPU starts on its queues ... */
			if (failed) break;
			for (oimgpu_lun *l : luns) oimgpu_lun_sync(l);			/* ... before any is waited for */
			for (uint32_t i = 2; i < vq.size(); i++) {
				if (vq[i].attached) signal_used(vq[i]);
			}
		}
	} else {
		for (uint32_t i = 2; i < vq.size(); i++) {
			if (vq[i].attached) signal_used(vq[i]);
		}
	}
}

/* ---- messages ------------------------------------------------------------------------------------ */

bool Session::handle_message()
{
	uint32_t hdr[3];
	std::vector<uint8_t> p;
	int fds[kMaxRegions], nfds = 0;
	for (int &f : fds) f = -1;
	const int rc = read_message(fd, hdr, p, fds, &nfds);
	if (rc <= 0 || hdr[0] >= REQ_MAX) {
		for (int i = 0; i < nfds; i++) close(fds[i]);
		return false;
	}
	const uint32_t req = hdr[0], flags = hdr[1];
	auto u64_of = [&]() { uint64_t v = 0; if (p.size() >= 8) memcpy(&v, p.data(), 8); return v; };
	auto state_of = [&](uint32_t *index, uint32_t *num) {
		*index = *num = 0;
		if (p.size() >= 8) { memcpy(index, p.data(), 4); memcpy(num, p.data() + 4, 4); }
	};
	auto close_fds = [&]() { for (int i = 0; i < nfds; i++) if (fds[i] >= 0) { close(fds[i]); fds[i] = -1; } };

	/* vhost_user_check_and_alloc_queue_pair (vhost_user.c:1053-1086) */
	{
		uint32_t idx = UINT32_MAX, num;
		if (req == SET_VRING_KICK || req == SET_VRING_CALL || req == SET_VRING_ERR) idx = (uint32_t)(u64_of() & kVringIdxMask);
		else if (req == SET_VRING_NUM || req == SET_VRING_BASE || req == SET_VRING_ENABLE || req == GET_VRING_BASE) state_of(&idx, &num);
		else if (req == SET_VRING_ADDR && p.size() >= 4) memcpy(&idx, p.data(), 4);
		if (idx != UINT32_MAX) {
			if (idx >= kMaxVring) { close_fds(); return false; }
			if (idx >= vq.size()) vq.resize(idx + 1);	/* the reference leaves holes below idx unallocated */
		}
	}

	int ret = 0;
	uint64_t v;
	uint32_t index, num;
	switch (req) {
	case GET_FEATURES:
		v = kFeatures;
		if (!reply(req, flags, &v, 8)) return false;
		break;
	case SET_FEATURES: {	/* vhost_user_set_features */
		v = u64_of();
		if (v & ~kFeaturesRef) break;	/* refused; the handler drops the status, so even a REPLY_ACK says 0 (vhost_user.c:1337-1339) */
		if (running && features != v) stop();
		features = v;
		break;
	}
	case GET_PROTOCOL_FEATURES:
		v = kProtocolFeatures;
		if (!reply(req, flags, &v, 8)) return false;
		break;
	case SET_PROTOCOL_FEATURES:	/* vhost_user_set_protocol_features */
		v = u64_of();
		if (v & ~kProtocolFeaturesRef) break;
		stop();
		protocol_features = v;
		break;
	case SET_OWNER:
		break;
	case RESET_OWNER:	/* vhost_user_reset_owner: back to the state of a fresh connection */
		stop();
		close_luns();
		free_mem();
		for (Vq &q : vq) { if (q.kickfd >= 0) close(q.kickfd); if (q.callfd >= 0) close(q.callfd); }
		vq.clear();
		features = protocol_features = 0;
		break;
	case SET_MEM_TABLE: {	/* vhost_user_set_mem_table: consumed by the next SET_VRING_ADDR */
		if (has_new_table) {
			for (uint32_t i = 0; i < kMaxRegions; i++) if (new_fds[i] >= 0) { close(new_fds[i]); new_fds[i] = -1; }
		}
		memset(&new_table, 0, sizeof(new_table));
		memcpy(&new_table, p.data(), p.size() < sizeof(new_table) ? p.size() : sizeof(new_table));
		if (new_table.nregions > kMaxRegions) new_table.nregions = kMaxRegions;
		for (uint32_t i = 0; i < kMaxRegions; i++) { new_fds[i] = fds[i]; fds[i] = -1; }
		has_new_table = true;
		break;
	}
	case SET_LOG_BASE:	/* dirty-page logging for live migration is not implemented; the reply is mandatory */
		close_fds();
		v = 0;
		if (!reply(req, flags, &v, 8)) return false;
		break;
	case SET_LOG_FD:
	case SET_VRING_ERR:
		close_fds();
		break;
	case SET_VRING_NUM:	/* vhost_user_set_vring_num */
		state_of(&index, &num);
		vq[index].size = num;
		break;
	case SET_VRING_ADDR: {	/* vhost_user_set_vring_addr */
		stop();
		if (has_new_table) { setup_mem(); has_new_table = false; }
		if (!have_mem || p.size() < 40) { ret = -1; break; }
		uint64_t a[4];	/* desc_user_addr, used_user_addr, avail_user_addr, log_guest_addr */
		memcpy(&index, p.data(), 4);
		memcpy(a, p.data() + 8, 32);
		Vq &q = vq[index];
		uint64_t len = 16ull * q.size;
		q.desc = qva_to_host(a[0], &len);
		if (!q.desc || len != 16ull * q.size) { q.desc = nullptr; ret = -1; break; }
		len = 4 + 2ull * q.size;
		q.avail = qva_to_host(a[2], &len);
		if (!q.avail || len != 4 + 2ull * q.size) { q.avail = nullptr; ret = -1; break; }
		len = 4 + 8ull * q.size;
		q.used = qva_to_host(a[1], &len);
		if (!q.used || len != 4 + 8ull * q.size) { q.used = nullptr; ret = -1; break; }
		const uint16_t used_idx = *(volatile uint16_t *)(q.used + 2);
		if (q.last_used != used_idx) q.last_used = q.last_avail = used_idx;	/* resume where the guest says we were */
		break;
	}
	case SET_VRING_BASE:	/* vhost_user_set_vring_base */
		stop();
		state_of(&index, &num);
		vq[index].last_used = vq[index].last_avail = (uint16_t)num;
		break;
	case GET_VRING_BASE: {	/* vhost_user_get_vring_base: stops the device */
		stop();
		state_of(&index, &num);
		Vq &q = vq[index];
		uint32_t st[2] = {index, q.last_used};
		if (q.kickfd >= 0) close(q.kickfd);
		q.kickfd = kFdUninit;
		if (q.callfd >= 0) close(q.callfd);
		q.callfd = kFdUninit;
		if (!reply(req, flags, st, 8)) return false;
		break;
	}
	case SET_VRING_KICK:
	case SET_VRING_CALL: {	/* vhost_user_set_vring_kick / _call */
		stop();
		v = u64_of();
		index = (uint32_t)(v & kVringIdxMask);
		int nf = (v & kVringNoFd) ? kFdInvalid : fds[0];
		if (!(v & kVringNoFd)) fds[0] = -1;
		int &slot = req == SET_VRING_KICK ? vq[index].kickfd : vq[index].callfd;
		if (slot >= 0) close(slot);
		slot = nf;
		break;
	}
	case GET_QUEUE_NUM:
		v = kMaxQueuePairs;
		if (!reply(req, flags, &v, 8)) return false;
		break;
	case SET_VRING_ENABLE:
		state_of(&index, &num);
		vq[index].enabled = num != 0;
		break;
	case GET_CONFIG: {
		/* the vhost-scsi backend has no get_config hook (vhost_scsi.c:126-136): the reference answers
		 * with the first 8 bytes of the request and size 8, which masters take as "not supported" */
		uint8_t b[8] = {};
		memcpy(b, p.data(), p.size() < 8 ? p.size() : 8);
		if (!reply(req, flags, b, 8)) return false;
		break;
	}
	case SET_CONFIG:
		ret = 1;
		break;
	case SEND_RARP:
	case NET_SET_MTU:
		break;
	default:
		ret = -1;
		break;
	}
	close_fds();
	if (flags & kNeedReply) {
		v = ret != 0;
		if (!reply(req, flags, &v, 8)) return false;
	}
	if (!running && is_ready()) start();
	return true;
}

void Session::teardown()
{
	VU_DEBUG("%s: session %d: teardown", srv->name.c_str(), fd);
	stop();
	close_luns();
	VU_DEBUG("%s: session %d: data path closed", srv->name.c_str(), fd);
	free_mem();
	VU_DEBUG("%s: session %d: guest memory released", srv->name.c_str(), fd);
	for (uint32_t i = 0; i < kMaxRegions; i++) if (has_new_table && new_fds[i] >= 0) close(new_fds[i]);
	for (Vq &q : vq) { if (q.kickfd >= 0) close(q.kickfd); if (q.callfd >= 0) close(q.callfd); }
	vq.clear();
	close(fd);
	fd = -1;
}

void Session::run()
{
	register_thread();
	for (int &f : new_fds) f = -1;
	while (!srv->stop.load()) {
		std::vector<pollfd> pf;
		pf.push_back({fd, POLLIN, 0});
		pf.push_back({wake, POLLIN, 0});
		if (running) {
			for (Vq &q : vq) if (q.kickfd >= 0) pf.push_back({q.kickfd, POLLIN, 0});
		}
		/* the management poller of the reference runs every 5 ms (MGMT_POLL_PERIOD_US); with a resident
		 * GPU poller this thread only relays interrupts and looks at the used indices much more often */
		const int timeout = !running ? -1 : polling ? 0 : 5;
		int n = poll(pf.data(), pf.size(), timeout);
		if (n < 0 && errno != EINTR) break;
		if (n > 0) {
			for (size_t i = 1; i < pf.size(); i++) {
				if (pf[i].revents & POLLIN) { eventfd_t c; eventfd_read(pf[i].fd, &c); }
			}
			if (pf[0].revents & (POLLIN | POLLHUP | POLLERR)) {
				if (!handle_message()) break;
			}
		}
		service();
		if (polling && n == 0) usleep(20);
	}
	teardown();
	unregister_thread();
	done.store(true);
}

void Server::accept_loop()
{
	while (!stop.load()) {
		pollfd pf[2] = {{lfd, POLLIN, 0}, {wake, POLLIN, 0}};
		if (poll(pf, 2, 500) <= 0) continue;
		if (!(pf[0].revents & POLLIN)) continue;
		int cfd = accept4(lfd, nullptr, nullptr, SOCK_CLOEXEC);
		if (cfd < 0) continue;
		auto s = std::make_unique<Session>();
		s->srv = this;
		s->fd = cfd;
		s->wake = eventfd(0, EFD_NONBLOCK | EFD_CLOEXEC);
		Session *sp = s.get();
		{
			std::lock_guard<std::mutex> lk(mu);
			/* reap sessions whose master has gone */
			for (size_t i = 0; i < sessions.size();) {
				if (sessions[i]->done.load()) {
					if (sessions[i]->th.joinable()) sessions[i]->th.join();
					close(sessions[i]->wake);
					sessions.erase(sessions.begin() + i);
				} else i++;
			}
			sessions.push_back(std::move(s));
		}
		sp->th = std::thread([sp] { sp->run(); });
	}
}

static void stop_server(Server &s)
{
	s.stop.store(true);
	eventfd_write(s.wake, 1);
	if (s.th.joinable()) s.th.join();
	{
		std::lock_guard<std::mutex> lk(s.mu);
		/* the sessions watch srv->stop; their `wake` eventfd gets them out of poll().  (Their sockets are theirs
		 * alone: touching ss->fd from here would race with a session that is closing it.) */
		for (auto &ss : s.sessions) eventfd_write(ss->wake, 1);
		for (auto &ss : s.sessions) {
			if (ss->th.joinable()) ss->th.join();
			close(ss->wake);
		}
		s.sessions.clear();
	}
	close(s.lfd);
	close(s.wake);
	unlink(s.path.c_str());
}

}  // namespace

void configure(const Config &cfg)
{
	g_cfg = cfg;
	signal(SIGUSR1, on_usr1);
	signal(SIGUSR2, on_usr2);
}

/* rte_vhost_driver_register + rte_vhost_driver_start (socket.c): the controller's listening socket */
int listen_ctrlr(const std::string &name, const std::string &path)
{
	sockaddr_un sa{};
	if (path.size() >= sizeof(sa.sun_path)) return -ENAMETOOLONG;
	std::lock_guard<std::mutex> lk(g_mu);
	if (g_servers.count(name)) return -EEXIST;
	int fd = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
	if (fd < 0) return -errno;
	/* bound under a temporary name, renamed once listening: the file's appearance is what deployments wait
	 * for (test/start-stop.make:20-23) */
	const std::string tmp = path + ".starting";
	if (tmp.size() >= sizeof(sa.sun_path)) { close(fd); return -ENAMETOOLONG; }
	sa.sun_family = AF_UNIX;
	snprintf(sa.sun_path, sizeof(sa.sun_path), "%s", tmp.c_str());
	unlink(tmp.c_str());
	if (bind(fd, (sockaddr *)&sa, sizeof(sa)) != 0 || listen(fd, 128) != 0 || rename(tmp.c_str(), path.c_str()) != 0) {
		int e = errno;
		close(fd);
		unlink(tmp.c_str());
		return -e;
	}
	auto s = std::make_unique<Server>();
	s->name = name;
	s->path = path;
	s->lfd = fd;
	s->wake = eventfd(0, EFD_NONBLOCK | EFD_CLOEXEC);
	Server *sp = s.get();
	s->th = std::thread([sp] { sp->accept_loop(); });
	g_servers[name] = std::move(s);
	return 0;
}

void close_ctrlr(const std::string &name)
{
	std::unique_ptr<Server> s;
	{
		std::lock_guard<std::mutex> lk(g_mu);
		auto it = g_servers.find(name);
		if (it == g_servers.end()) return;
		s = std::move(it->second);
		g_servers.erase(it);
	}
	stop_server(*s);
}

void notify_target(const std::string &name, int scsi_target_num, bool added)
{
	std::lock_guard<std::mutex> lk(g_mu);
	auto it = g_servers.find(name);
	if (it == g_servers.end()) return;
	std::lock_guard<std::mutex> lk2(it->second->mu);
	for (auto &ss : it->second->sessions) {
		if (ss->done.load()) continue;
		{
			std::lock_guard<std::mutex> lk3(ss->ev_mu);
			ss->events.push_back({scsi_target_num, added});
		}
		eventfd_write(ss->wake, 1);
	}
}

int active_sessions(const std::string &name)
{
	std::lock_guard<std::mutex> lk(g_mu);
	auto it = g_servers.find(name);
	if (it == g_servers.end()) return 0;
	std::lock_guard<std::mutex> lk2(it->second->mu);
	int n = 0;
	for (auto &ss : it->second->sessions) n += !ss->done.load();	/* every connected master counts (new_connection inserts into vdev->vsessions, vhost.c) */
	return n;
}

void shutdown()
{
	std::map<std::string, std::unique_ptr<Server>> all;
	{
		std::lock_guard<std::mutex> lk(g_mu);
		all.swap(g_servers);
	}
	for (auto &kv : all) stop_server(*kv.second);
}

}  // namespace vhost_user
