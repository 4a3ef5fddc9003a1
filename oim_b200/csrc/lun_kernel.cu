/*
 * lun_kernel.cu — oim_lun_queue_kernel: see lun_kernel.cuh for the map onto the reference.
 *
 * Launch shape: grid = min(#queues, SMs x CTAs/SM), block = 256 threads.  CTA b owns queues
 * b, b+grid, ... and for each runs passes of <= 32 requests:
 *   1. stage the 32 request slots (2 KiB) in shared memory, 16 B per thread, coalesced
 *   2. warp 0: lane i parses request i (SG walk, LUN check, CDB decode, limits), the warp finds
 *      LBA hazards, and the lanes emit SG segments + completion records to shared memory
 *   3. all warps move payload, one 4 KiB unit per warp-step, 8 x 16-byte loads in flight per lane
 *   4. warp 0 writes the 32 completion records (48 B each) coalesced
 */
#include "lun_kernel.cuh"

namespace oimgpu {

/* ---- parser: one lane, one request ------------------------------------------------------- */

__device__ __forceinline__ uint32_t units_of(uint64_t len) { return (uint32_t)((len + kUnitBytes - 1) / kUnitBytes); }

/* spdk_bdev_bytes_to_blocks + spdk_bdev_io_valid_blocks (bdev.c:2474-2509) on (offset, nbytes) */
__device__ __forceinline__ bool bdev_range_ok(const LunCtx &L, uint64_t lba, uint64_t nbytes, uint64_t *nblk)
{
	uint64_t nb;
	if (L.block_shift != 0xffffffffu) {
		nb = nbytes >> L.block_shift;
		if ((nb << L.block_shift) != nbytes) return false;
	} else {
		nb = nbytes / L.block_size;
		if (nb * L.block_size != nbytes) return false;
	}
	if (lba + nb < lba) return false;
	if (lba + nb > L.num_blocks) return false;
	*nblk = nb;
	return true;
}

/* spdk_bdev_scsi_readwrite + _read/_write (scsi_bdev.c:1456-1511, 1318-1411) */
__device__ __forceinline__ void scsi_readwrite(const LunCtx &L, LaneState &s, uint32_t dxfer_dir, uint32_t transfer_len,
					       uint64_t lba, uint32_t xfer_len, bool is_read)
{
	s.data_transferred = 0;
	if (dxfer_dir != OIMGPU_DIR_NONE && dxfer_dir != (is_read ? OIMGPU_DIR_FROM_DEV : OIMGPU_DIR_TO_DEV)) {
		set_check(s, SK_NO_SENSE, ASC_NONE);
		return;
	}
	if (L.num_blocks <= lba || L.num_blocks - lba < xfer_len) {
		set_check(s, SK_ILLEGAL_REQUEST, ASC_LBA_OOR);
		return;
	}
	if (xfer_len == 0) {
		s.status = SC_GOOD;
		return;
	}
	if (xfer_len > OIMGPU_MAX_XFER_BYTES / L.block_size) {
		set_check(s, SK_ILLEGAL_REQUEST, ASC_INVALID_FIELD);
		return;
	}
	uint64_t nblk;
	if (!is_read) {
		if ((uint64_t)xfer_len * L.block_size > transfer_len) {
			set_check(s, SK_NO_SENSE, ASC_NONE);
			return;
		}
	}
	/* both directions move task->length bytes, not xfer_len blocks (scsi_bdev.c:1333, 1387) */
	if (!bdev_range_ok(L, lba, s.length, &nblk)) {
		set_check(s, SK_NO_SENSE, ASC_NONE);
		return;
	}
	s.data_transferred = s.length;
	s.off = lba * L.block_size;
	s.store_lo = lba;
	s.store_hi = lba + nblk;
	s.op = is_read ? OP_READ : OP_WRITE;
	s.hazard = nblk ? (is_read ? 1 : 2) : 0;
}

/* UNMAP parameter list walk (scsi_bdev.c:1545-1679).  emit == nullptr: validate and count;
 * otherwise also write one zero-fill segment per accepted descriptor. */
__device__ inline void scsi_unmap(const LunCtx &L, const QueueDesc &q, const oimgpu_req &r, uint32_t iovcnt,
				  LaneState &s, Segment *emit, uint32_t *emit_unit, uint16_t wave)
{
	const uint32_t data_len = s.length;
	int desc_count = -1;
	if (data_len >= 8) {
		uint16_t ddl = (uint16_t)(gather_byte(q, r, iovcnt, 2) << 8 | gather_byte(q, r, iovcnt, 3));
		if (ddl <= data_len - 8 && ddl / 16 <= OIMGPU_MAX_UNMAP_DESC) desc_count = ddl / 16;
	}
	if (desc_count < 0) {
		if (!emit) set_check(s, SK_ILLEGAL_REQUEST, ASC_INVALID_FIELD);
		return;
	}
	uint32_t nseg = 0, units = 0;
	for (int i = 0; i < desc_count; i++) {
		uint64_t ob = 0;
		uint32_t nb = 0;
		for (int k = 0; k < 8; k++) ob = ob << 8 | gather_byte(q, r, iovcnt, 8 + 16 * i + k);
		for (int k = 8; k < 12; k++) nb = nb << 8 | gather_byte(q, r, iovcnt, 8 + 16 * i + k);
		if (nb == 0) continue;
		if (ob + nb < ob || ob + nb > L.num_blocks) {
			/* spdk_bdev_unmap_blocks -> -EINVAL: earlier descriptors stay applied, the rest are skipped */
			if (!emit) set_check(s, SK_NO_SENSE, ASC_NONE);
			break;
		}
		uint64_t bytes = (uint64_t)nb * L.block_size;
		if (emit) {
			Segment &g = emit[nseg];
			g.src = nullptr;
			g.dst = L.store[0] + ob * L.block_size;
			g.len = bytes;
			g.first_unit = *emit_unit;
			g.wave = wave;
			g.mirror = 1;
			*emit_unit += units_of(bytes);
		}
		nseg++;
		units += units_of(bytes);
	}
	if (!emit) {
		s.nseg = nseg;
		s.units = units;
		s.op = OP_UNMAP;
		s.hazard = nseg ? 3 : 0;
	}
}

__device__ inline void parse_request(const LunCtx &L, const QueueDesc &q, const oimgpu_req &r, LaneState &s)
{
	const uint32_t cnt = r.iovcnt;
	const bool from_dev = (r.dir == OIMGPU_DIR_FROM_DEV) || cnt == 0;
	const uint32_t dxfer_dir = from_dev ? OIMGPU_DIR_FROM_DEV : OIMGPU_DIR_TO_DEV;
	uint32_t len = 0, nonzero = 0, units = 0;

	s.op = OP_NONE; s.nseg = 0; s.valid = 1; s.length = 0; s.off = 0;
	s.store_lo = s.store_hi = 0;
	s.status = SC_GOOD; s.sk = 0; s.asc = 0;
	s.data_transferred = 0; s.units = 0; s.hazard = 0; s.response = OIMGPU_S_OK; s.resp_valid = 1;

	/* ---- task_data_setup (vhost_scsi.c:490-624): walk the SG list ---- */
	for (uint32_t j = 0; j < cnt; j++) {
		if (j >= OIMGPU_IOVS_MAX) { s.valid = 0; break; }
		const oimgpu_iov v = q.iovs[(r.iov_start + j) & q.iov_mask];
		if (v.addr == 0) { s.valid = 0; break; }
		len += v.len;
		if (v.len) { nonzero++; units += units_of(v.len); }
	}
	if (!s.valid) {
		s.used_len = 0;		/* invalid_request(): used element only (vhost_scsi.c:347-358) */
		s.resp_valid = 0;
		s.response = 0;
		return;
	}
	s.length = len;
	s.used_len = from_dev ? OIMGPU_RESP_SIZE + len : OIMGPU_RESP_SIZE;

	/* ---- spdk_vhost_scsi_task_init_target (vhost_scsi.c:361-387) ---- */
	const uint16_t lun_id = (uint16_t)((((uint16_t)r.lun[2] << 8) | r.lun[3]) & 0x3FFF);
	if (r.lun[0] != 1 || r.lun[1] >= OIMGPU_CTRLR_MAX_DEVS || (r.lun[1] != L.target) ) {
		s.response = OIMGPU_S_BAD_TARGET;	/* resp->response only; nothing else is written */
		return;
	}
	const bool null_lun = L.removed || lun_id != 0;
	const uint8_t *cdb = r.cdb;

	if (null_lun) {
		/* spdk_scsi_task_process_null_lun (task.c:258-293) */
		if (cdb[0] == 0x12) {
			uint8_t *buf = s.scratch;
			for (int k = 0; k < 36; k++) buf[k] = 0;
			buf[0] = 0x03 << 5 | 0x1f;
			buf[4] = 36 - 5;
			uint32_t alloc_len = be16(&cdb[3]);
			if (scatter_small(q, r, cnt, len, buf, alloc_len < 36 ? alloc_len : 36, s) >= 0) {
				s.data_transferred = 36;
				s.status = SC_GOOD;
			}
		} else {
			set_check(s, SK_ILLEGAL_REQUEST, ASC_LUN_NOT_SUPPORTED);
			s.data_transferred = 0;
		}
		return;
	}
	if (L.lun_removed) {
		set_check(s, SK_ABORTED_COMMAND, ASC_NONE);	/* spdk_scsi_task_process_abort */
		return;
	}

	/* ---- spdk_bdev_scsi_process_block (scsi_bdev.c:1681-1802) ---- */
	switch (cdb[0]) {
	case 0x08: case 0x0a: {
		uint64_t lba = (uint64_t)cdb[1] << 16 | (uint64_t)cdb[2] << 8 | cdb[3];
		uint32_t xl = cdb[4] ? cdb[4] : 256;
		scsi_readwrite(L, s, dxfer_dir, len, lba, xl, cdb[0] == 0x08);
		break;
	}
	case 0x28: case 0x2a:
		scsi_readwrite(L, s, dxfer_dir, len, be32(&cdb[2]), be16(&cdb[7]), cdb[0] == 0x28);
		break;
	case 0xa8: case 0xaa:
		scsi_readwrite(L, s, dxfer_dir, len, be32(&cdb[2]), be32(&cdb[6]), cdb[0] == 0xa8);
		break;
	case 0x88: case 0x8a:
		scsi_readwrite(L, s, dxfer_dir, len, be64(&cdb[2]), be32(&cdb[10]), cdb[0] == 0x88);
		break;
	case 0x25: {	/* READ CAPACITY (10) */
		uint8_t *buf = s.scratch;
		uint64_t last = L.num_blocks - 1;
		uint32_t v = last > 0xffffffffULL ? 0xffffffffu : (uint32_t)last;
		buf[0] = v >> 24; buf[1] = v >> 16; buf[2] = v >> 8; buf[3] = v;
		buf[4] = L.block_size >> 24; buf[5] = L.block_size >> 16; buf[6] = L.block_size >> 8; buf[7] = L.block_size;
		uint32_t l = len < 8 ? len : 8;
		if (scatter_small(q, r, cnt, len, buf, l, s) >= 0) {
			s.data_transferred = l;
			s.status = SC_GOOD;
		}
		break;
	}
	case 0x9e:
		if ((cdb[1] & 0x1f) == 0x10) {	/* READ CAPACITY (16) */
			uint8_t *buf = s.scratch;
			for (int k = 0; k < 32; k++) buf[k] = 0;
			uint64_t last = L.num_blocks - 1;
			for (int k = 0; k < 8; k++) buf[k] = (uint8_t)(last >> (56 - 8 * k));
			buf[8] = L.block_size >> 24; buf[9] = L.block_size >> 16; buf[10] = L.block_size >> 8; buf[11] = L.block_size;
			buf[14] |= 1 << 7;
			uint32_t al = be32(&cdb[10]);
			uint32_t l = al < 32 ? al : 32;
			if (scatter_small(q, r, cnt, len, buf, l, s) >= 0) {
				s.data_transferred = l;
				s.status = SC_GOOD;
			}
		} else {
			set_check(s, SK_ILLEGAL_REQUEST, ASC_INVALID_OPCODE);
		}
		break;
	case 0x35: case 0x91: {	/* SYNCHRONIZE CACHE: bounds check only; FLUSH is a no-op on a RAM disk */
		uint64_t lba; uint32_t n;
		if (cdb[0] == 0x35) { lba = be32(&cdb[2]); n = be16(&cdb[7]); }
		else { lba = be64(&cdb[2]); n = be32(&cdb[10]); }
		if (n == 0) n = (uint32_t)(L.num_blocks - lba);
		if (n != 0 && (lba >= L.num_blocks || n > L.num_blocks || lba > L.num_blocks - n)) {
			set_check(s, SK_NO_SENSE, ASC_NONE);
		}
		break;
	}
	case 0x42:
		scsi_unmap(L, q, r, cnt, s, nullptr, nullptr, 0);
		break;
	/* ---- spdk_bdev_scsi_process_primary (scsi_bdev.c:1827-2077), table-free commands ---- */
	case 0x03:	/* REQUEST SENSE */
		if (!(cdb[1] & 0x1)) {
			uint8_t *buf = s.scratch;
			for (int k = 0; k < 18; k++) buf[k] = 0;
			buf[0] = 0xf0; buf[7] = 10;
			uint32_t al = cdb[4];
			scatter_small(q, r, cnt, len, buf, al < 18 ? al : 18, s);
			s.data_transferred = al < 18 ? al : 18;
		}
		s.status = SC_GOOD;	/* rc >= 0 path overrides whatever status was set (scsi_bdev.c:2066-2069) */
		break;
	case 0x4c: case 0x4d:	/* LOG SELECT / LOG SENSE */
		set_check(s, SK_ILLEGAL_REQUEST, ASC_INVALID_OPCODE);
		break;
	case 0x00: case 0x1b:	/* TEST UNIT READY / START STOP UNIT */
		break;
	default:
		set_check(s, SK_ILLEGAL_REQUEST, ASC_INVALID_OPCODE);
		break;
	}
	if (s.op == OP_READ || s.op == OP_WRITE) {
		s.nseg = nonzero;
		s.units = units;
	}
}

/* emit the payload segments of one parsed request into shared memory */
__device__ inline void emit_segments(const LunCtx &L, const QueueDesc &q, const oimgpu_req &r, LaneState &s,
				     Segment *out, uint32_t first_unit, uint16_t wave)
{
	if (s.op == OP_UNMAP) {
		uint32_t u = first_unit;
		scsi_unmap(L, q, r, r.iovcnt, s, out, &u, wave);
		return;
	}
	if (s.op != OP_READ && s.op != OP_WRITE) return;
	uint8_t *pos = L.store[0] + s.off;
	uint32_t k = 0, u = first_unit;
	for (uint32_t j = 0; j < r.iovcnt; j++) {
		const oimgpu_iov v = q.iovs[(r.iov_start + j) & q.iov_mask];
		if (v.len == 0) continue;
		Segment &g = out[k++];
		uint8_t *client = (uint8_t *)(uintptr_t)v.addr;
		if (s.op == OP_READ) { g.src = pos; g.dst = client; g.mirror = 0; }
		else { g.src = client; g.dst = pos; g.mirror = 1; }
		g.len = v.len;
		g.first_unit = u;
		g.wave = wave;
		u += units_of(v.len);
		pos += v.len;
	}
}

/* ---- the kernel ----------------------------------------------------------------------------- */

#ifndef OIM_MIN_BLOCKS
#define OIM_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(kThreads, OIM_MIN_BLOCKS)
oim_lun_queue_kernel(LunCtx *lun, const QueueDesc *queues, uint32_t nqueues)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	PassShared &sh = *reinterpret_cast<PassShared *>(smem_raw);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const LunCtx &L = *lun;

	for (uint32_t qi = blockIdx.x; qi < nqueues; qi += gridDim.x) {
		const QueueDesc q = queues[qi];
		for (uint32_t done = 0; done < q.count; done += kPass) {
			const uint32_t n = min((uint32_t)kPass, q.count - done);
			const uint32_t slot0 = q.head + done;

			/* 1. stage request slots: n x 64 B as 16-byte vectors */
			for (uint32_t v = tid; v < n * 4; v += kThreads) {
				const uint32_t slot = (slot0 + (v >> 2)) & q.ring_mask;
				reinterpret_cast<int4 *>(&sh.req[v >> 2])[v & 3] =
					ld_cg16(reinterpret_cast<const int4 *>(&q.reqs[slot]) + (v & 3));
			}
			__syncthreads();

			/* 2. parse + hazards + first round bookkeeping (warp 0) */
			LaneState &s = sh.lane[lane];	/* meaningful for warp 0 only */
			uint16_t wave = 0;
			uint32_t nwaves = 1;
			if (warp == 0) {
				const bool active = (uint32_t)lane < n;
				if (active) parse_request(L, q, sh.req[lane], s);
				else { s.nseg = 0; s.units = 0; s.hazard = 0; s.op = OP_NONE; s.valid = 0; s.store_lo = s.store_hi = 0; }

				/* hazard waves.  Writers are visited in ring order; a writer is pushed behind every
				 * earlier request it overlaps, every later overlapping request behind the writer. */
				uint32_t writers = __ballot_sync(0xffffffffu, active && s.hazard >= 2);
				uint32_t touching = __ballot_sync(0xffffffffu, active && s.hazard != 0);
				if (writers && (touching & (touching - 1))) {
					uint32_t w = writers;
					while (w) {
						const int j = __ffs(w) - 1;
						w &= w - 1;
						const uint64_t jlo = __shfl_sync(0xffffffffu, s.store_lo, j);
						const uint64_t jhi = __shfl_sync(0xffffffffu, s.store_hi, j);
						const int jhaz = __shfl_sync(0xffffffffu, (int)s.hazard, j);
						const bool overlap = s.hazard != 0 && lane != j &&
							(jhaz == 3 || s.hazard == 3 || (s.store_lo < jhi && jlo < s.store_hi));
						/* earlier overlapping requests decide the writer's wave */
						uint32_t need = (overlap && lane < j) ? (uint32_t)wave + 1 : 0;
						need = __reduce_max_sync(0xffffffffu, need);
						if (lane == j && need > wave) wave = (uint16_t)need;
						const uint32_t jw = __shfl_sync(0xffffffffu, (uint32_t)wave, j);
						if (overlap && lane > j && jw + 1 > wave) wave = (uint16_t)(jw + 1);
					}
					nwaves = __reduce_max_sync(0xffffffffu, (uint32_t)wave) + 1;
				}
				if (lane == 0) sh.nwaves = nwaves;
			}

			/* rounds: as many whole requests as fit in the segment table */
			uint32_t r0 = 0;
			while (r0 < n) {
				if (warp == 0) {
					/* prefix sums of segment and unit counts over requests r0.. */
					const bool in = (uint32_t)lane >= r0 && (uint32_t)lane < n;
					uint32_t segs = in ? s.nseg : 0, units = in ? s.units : 0;
					uint32_t seg_incl = segs, unit_incl = units;
#pragma unroll
					for (int o = 1; o < 32; o <<= 1) {
						uint32_t a = __shfl_up_sync(0xffffffffu, seg_incl, o);
						uint32_t b = __shfl_up_sync(0xffffffffu, unit_incl, o);
						if (lane >= o) { seg_incl += a; unit_incl += b; }
					}
					const uint32_t fits = __ballot_sync(0xffffffffu, in && seg_incl <= (uint32_t)kSegCap);
					/* lanes r0..r1-1 fit (prefix property: seg_incl is monotonic) */
					const uint32_t r1 = r0 + __popc(fits);
					const bool mine = in && (uint32_t)lane < r1;
					if (mine && segs) {
						emit_segments(L, q, sh.req[lane], s, &sh.seg[seg_incl - segs], unit_incl - units, wave);
					}
					const uint32_t last = r1 - 1;
					const uint32_t tot_seg = __shfl_sync(0xffffffffu, seg_incl, last);
					const uint32_t tot_unit = __shfl_sync(0xffffffffu, unit_incl, last);
					if (lane == 0) { sh.nseg = tot_seg; sh.nunits = tot_unit; sh.round_reqs = r1 - r0; }
				}
				__syncthreads();

				/* 3. move payload */
				const uint32_t nseg = sh.nseg, nunits = sh.nunits, nw = sh.nwaves;
				for (uint32_t w = 0; w < nw; w++) {
					for (uint32_t u = warp; u < nunits; u += kWarps) {
						uint32_t lo = 0, hi = nseg;	/* last segment with first_unit <= u */
						while (hi - lo > 1) {
							const uint32_t mid = (lo + hi) >> 1;
							if (sh.seg[mid].first_unit <= u) lo = mid; else hi = mid;
						}
						const Segment &g = sh.seg[lo];
						if (nw > 1 && g.wave != w) continue;
						const uint64_t off = (uint64_t)(u - g.first_unit) * kUnitBytes;
						const uint32_t nbytes = (uint32_t)min((uint64_t)kUnitBytes, g.len - off);
						if (g.src) move_unit(g.dst + off, g.src + off, nbytes, lane);
						else zero_unit(g.dst + off, nbytes, lane);
						if (g.mirror && L.nreplicas > 1) {
							/* mirrored bdev: same bytes to every peer replica over NVLink (P2P stores) */
							const uint64_t soff = (uint64_t)(g.dst - L.store[0]) + off;
							for (uint32_t rep = 1; rep < L.nreplicas; rep++) {
								if (g.src) move_unit(L.store[rep] + soff, g.src + off, nbytes, lane);
								else zero_unit(L.store[rep] + soff, nbytes, lane);
							}
						}
					}
					if (nw > 1) __syncthreads();
				}
				r0 += sh.round_reqs;
				__syncthreads();
			}

			/* 4. completion records (spdk_vhost_scsi_task_cpl, vhost_scsi.c:311-331) */
			if (warp == 0) {
				if ((uint32_t)lane < n) {
					__align__(16) oimgpu_cpl c;
					int4 *cz = reinterpret_cast<int4 *>(&c);
					cz[0] = cz[1] = cz[2] = make_int4(0, 0, 0, 0);
					c.tag = sh.req[lane].tag;
					c.used_len = s.used_len;
					c.resp_valid = s.resp_valid;
					c.response = s.response;
					if (s.resp_valid && s.response == OIMGPU_S_OK) {
						c.status = s.status;
						if (s.status != SC_GOOD) {
							c.sense[0] = 0xf0; c.sense[2] = s.sk & 0xf; c.sense[7] = 10;
							c.sense[12] = s.asc; c.sense[13] = 0;
							c.sense_len = OIMGPU_SENSE_SIZE;
						}
						c.resid = s.length - s.data_transferred;
						c.data_transferred = s.data_transferred;
					}
					sh.cpl[lane] = c;
				}
				__syncwarp();
				for (uint32_t v = lane; v < n * 3; v += 32) {
					const uint32_t slot = (slot0 + v / 3) & q.ring_mask;
					st_cg16(reinterpret_cast<int4 *>(&q.cpls[slot]) + v % 3,
						reinterpret_cast<const int4 *>(&sh.cpl[v / 3])[v % 3]);
				}
				/* counters for get_bdevs_iostat: one atomic per counter per pass */
				const bool ok = (uint32_t)lane < n && s.resp_valid && s.response == OIMGPU_S_OK && s.status == SC_GOOD;
				const uint32_t rd = __popc(__ballot_sync(0xffffffffu, ok && s.op == OP_READ));
				const uint32_t wr = __popc(__ballot_sync(0xffffffffu, ok && s.op == OP_WRITE));
				const uint32_t um = __popc(__ballot_sync(0xffffffffu, ok && s.op == OP_UNMAP));
				const uint32_t er = __popc(__ballot_sync(0xffffffffu, (uint32_t)lane < n && !ok));
				uint32_t rb = (ok && s.op == OP_READ) ? s.length : 0;
				uint32_t wb = (ok && s.op == OP_WRITE) ? s.length : 0;
				unsigned long long rbt = __reduce_add_sync(0xffffffffu, rb >> 9), wbt = __reduce_add_sync(0xffffffffu, wb >> 9);
				if (lane == 0) {
					if (rd) { atomicAdd(&lun->stats[0], rd); atomicAdd(&lun->stats[4], rbt << 9); }
					if (wr) { atomicAdd(&lun->stats[1], wr); atomicAdd(&lun->stats[5], wbt << 9); }
					if (um) atomicAdd(&lun->stats[2], um);
					if (er) atomicAdd(&lun->stats[7], er);
				}
			}
			__syncthreads();
		}
	}
}

/* struct spdk_copy_engine.copy / .fill (S/include/spdk_internal/copy_engine.h:47-53) for
 * device-resident buffers: a plain grid-stride mover built from the same unit routines */
__global__ void __launch_bounds__(kThreads)
oim_copy_kernel(uint8_t *dst, const uint8_t *src, uint64_t nbytes)
{
	const int lane = threadIdx.x & 31;
	const uint64_t warp = (uint64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
	const uint64_t nwarps = (uint64_t)gridDim.x * kWarps;
	const uint64_t nunits = (nbytes + kUnitBytes - 1) / kUnitBytes;
	for (uint64_t u = warp; u < nunits; u += nwarps) {
		const uint64_t off = u * kUnitBytes;
		const uint32_t n = (uint32_t)min((uint64_t)kUnitBytes, nbytes - off);
		if (src) move_unit(dst + off, src + off, n, lane);
		else zero_unit(dst + off, n, lane);
	}
}

__global__ void __launch_bounds__(kThreads)
oim_fill_kernel(uint8_t *dst, uint8_t fill, uint64_t nbytes)
{
	const uint64_t i0 = ((uint64_t)blockIdx.x * kThreads + threadIdx.x) * 16;
	const uint64_t stride = (uint64_t)gridDim.x * kThreads * 16;
	const uint32_t w = fill * 0x01010101u;
	const int4 v = make_int4(w, w, w, w);
	if (((uintptr_t)dst & 15) == 0) {
		for (uint64_t i = i0; i + 16 <= nbytes; i += stride) st_cg16(dst + i, v);
		const uint64_t tail = nbytes & ~15ull;
		if (blockIdx.x == 0 && threadIdx.x < (nbytes & 15)) dst[tail + threadIdx.x] = fill;
	} else {
		for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < nbytes; i += (uint64_t)gridDim.x * kThreads) dst[i] = fill;
	}
}

size_t lun_kernel_smem_bytes() { return sizeof(PassShared); }

}  // namespace oimgpu
