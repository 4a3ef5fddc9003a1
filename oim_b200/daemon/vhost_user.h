/*
 * vhost_user.h — the vhost-user slave of oim-gpu-vhost: what lets QEMU's vhost-user-scsi-pci connect to
 * <socket-dir>/<controller> exactly as it connects to SPDK's `vhost`.
 *
 * Replaces, for the vhost-scsi case, S/lib/vhost/rte_vhost/{socket,vhost_user,vhost,fd_man}.c (the
 * protocol) and the session start/stop half of S/lib/vhost/vhost.c:1044-1134 + vhost_scsi.c:1236-1400.
 * The request queues are handed to the GPU (oimgpu_vq_attach); the control and event queues, which carry
 * a handful of messages per VM lifetime, are served here on the CPU.
 */
#pragma once

#include <string>

namespace vhost_user {

struct Config {
	bool control_only = false;	/* no CUDA: handshake only (protocol tests) */
	bool poller = false;		/* resident GPU poller per session instead of one launch per kick */
	bool spread = true;		/* several GPUs: a session's request queues are dealt out among all of them */
};

void configure(const Config &cfg);
/* construct_vhost_scsi_controller: create and listen on `path` (rte_vhost_driver_register + _start) */
int listen_ctrlr(const std::string &name, const std::string &path);
/* remove_vhost_controller: stop listening, drop the socket file */
void close_ctrlr(const std::string &name);
/* add_vhost_scsi_lun / remove_vhost_scsi_target on a controller with live sessions: VIRTIO_SCSI_T_TRANSPORT_RESET
 * events on their event queues (vhost_scsi.c:939-946, 1059-1062) */
void notify_target(const std::string &name, int scsi_target_num, bool added);
/* number of sessions of the controller whose device is started (get_vhost_controllers could report it) */
int active_sessions(const std::string &name);
void shutdown();

}  // namespace vhost_user
