// Micro-benchmark (evidence for DESIGN.md): which SM-originated store reaches PEER HBM over NVLink fastest?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/pcie_store_bench tools/pcie_store_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ int4 ld16(const void *p) { int4 r; asm volatile("ld.global.cg.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x),"=r"(r.y),"=r"(r.z),"=r"(r.w) : "l"(p)); return r; }
__device__ __forceinline__ void st16cg(void *p, int4 v) { asm volatile("st.global.cg.v4.s32 [%0], {%1,%2,%3,%4};" :: "l"(p),"r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w) : "memory"); }
__device__ __forceinline__ void st16wt(void *p, int4 v) { asm volatile("st.global.wt.v4.s32 [%0], {%1,%2,%3,%4};" :: "l"(p),"r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w) : "memory"); }
__device__ __forceinline__ void st16cs(void *p, int4 v) { asm volatile("st.global.cs.v4.s32 [%0], {%1,%2,%3,%4};" :: "l"(p),"r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w) : "memory"); }
__device__ __forceinline__ void st32(void *p, int4 a, int4 b) { asm volatile("st.global.v8.s32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p),"r"(a.x),"r"(a.y),"r"(a.z),"r"(a.w),"r"(b.x),"r"(b.y),"r"(b.z),"r"(b.w) : "memory"); }

template <int MODE> __global__ void copy_k(uint8_t *dst, const uint8_t *src, size_t n)
{
	const int lane = threadIdx.x & 31;
	size_t warp = (size_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, nw = (size_t)gridDim.x * (blockDim.x / 32);
	for (size_t u = warp; u < n / 4096; u += nw) {
		const uint8_t *s = src + u * 4096; uint8_t *d = dst + u * 4096;
		if (MODE == 3) {
			int4 a[4], b[4];
			for (int k = 0; k < 4; k++) { a[k] = ld16(s + (lane + 32 * k) * 32); b[k] = ld16(s + (lane + 32 * k) * 32 + 16); }
			for (int k = 0; k < 4; k++) st32(d + (lane + 32 * k) * 32, a[k], b[k]);
		} else {
			int4 r[8];
			for (int k = 0; k < 8; k++) r[k] = ld16(s + (lane + 32 * k) * 16);
			for (int k = 0; k < 8; k++) { void *p = d + (lane + 32 * k) * 16; if (MODE == 0) st16cg(p, r[k]); else if (MODE == 1) st16wt(p, r[k]); else st16cs(p, r[k]); }
		}
	}
}

// TMA bulk: one lane per warp moves 4 KiB HBM -> smem -> host with cp.async.bulk
__global__ void copy_tma(uint8_t *dst, const uint8_t *src, size_t n)
{
	extern __shared__ __align__(128) uint8_t sm[];
	__shared__ uint64_t bar[8];
	const int w = threadIdx.x / 32, lane = threadIdx.x & 31;
	uint8_t *buf = sm + w * 4096;
	uint32_t sbuf = (uint32_t)__cvta_generic_to_shared(buf), sbar = (uint32_t)__cvta_generic_to_shared(&bar[w]);
	if (lane == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(sbar)); asm volatile("fence.mbarrier_init.release.cluster;"); }
	__syncwarp();
	size_t warp = (size_t)blockIdx.x * (blockDim.x / 32) + w, nw = (size_t)gridDim.x * (blockDim.x / 32);
	uint32_t phase = 0;
	for (size_t u = warp; u < n / 4096; u += nw) {
		if (lane == 0) {
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 4096;" :: "r"(sbar) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 4096, [%2];" :: "r"(sbuf), "l"(src + u * 4096), "r"(sbar) : "memory");
			asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" :: "r"(sbar), "r"(phase) : "memory");
			phase ^= 1;
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4096;" :: "l"(dst + u * 4096), "r"(sbuf) : "memory");
			asm volatile("cp.async.bulk.commit_group;" ::: "memory");
			asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
		}
		__syncwarp();
	}
	if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main()
{
	size_t n = 1ull << 30;
	uint8_t *d_src, *h_dst;
	cudaMalloc(&d_src, n); cudaMemset(d_src, 7, n);
	cudaSetDevice(1); cudaMalloc(&h_dst, n); cudaSetDevice(0); cudaDeviceEnablePeerAccess(1, 0);
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	const char *names[] = {"st.cg.v4", "st.wt.v4", "st.cs.v4", "st.v8(32B)", "tma bulk 4KiB"};
	for (int grid : {148, 296, 592, 1184}) for (int mode = 0; mode < 5; mode++) {
		float best = 1e9;
		for (int it = 0; it < 3; it++) {
			cudaEventRecord(a);
			if (mode == 0) copy_k<0><<<grid, 256>>>(h_dst, d_src, n);
			else if (mode == 1) copy_k<1><<<grid, 256>>>(h_dst, d_src, n);
			else if (mode == 2) copy_k<2><<<grid, 256>>>(h_dst, d_src, n);
			else if (mode == 3) copy_k<3><<<grid, 256>>>(h_dst, d_src, n);
			else copy_tma<<<grid, 256, 8 * 4096>>>(h_dst, d_src, n);
			cudaEventRecord(b); cudaEventSynchronize(b);
			float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
		}
		cudaError_t e = cudaGetLastError();
		printf("grid %4d %-14s %6.2f GB/s %s\n", grid, names[mode], n / (best * 1e-3) / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
	}
	// host -> device direction (kernel loads from pinned host): the WRITE path of the LUN
	uint8_t *d_dst; cudaMalloc(&d_dst, n);
	for (int grid : {296, 1184}) {
		float best = 1e9;
		for (int it = 0; it < 3; it++) { cudaEventRecord(a); copy_k<0><<<grid, 256>>>(d_dst, h_dst, n); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
		printf("grid %4d peer->local ld.cg.v4 x8 %6.2f GB/s\n", grid, n / (best * 1e-3) / 1e9);
	}
	return 0;
}
