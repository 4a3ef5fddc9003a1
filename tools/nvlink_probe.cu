// nvlink_probe — what does one B200 reach when it replicates a written extent to its peers?  (evidence for the
// mirrored-bdev design; numbers go into bench.py's mirror leg as the NVLink denominator)
//
//   ce        cudaMemcpyPeerAsync 0 -> 1 (the copy engine: what a separate "send" step does at best)
//   st        SM-originated st.global.cg.v4 from GPU 0 into GPU 1's HBM (what the fused mirror kernel does)
//   st2       one load, two stores: local HBM + peer HBM (the fused mirror's exact shape)
//   mc        multimem.st on a cuMulticast object bound on ALL visible GPUs: one store, the NVSwitch replicates
//   pull      GPU 1 loads GPU 0's HBM with ld.global.cg.v4 (the replica pulls)
//   ce_fan    copy engine 0 -> every other GPU at once (R-1 peer copies on R-1 streams)
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/nvlink_probe.bin tools/nvlink_probe.cu -lcuda
// run:   gpurun --gpus 2 -- tools/nvlink_probe.bin       (or --gpus 4 for the R=4 multicast row)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s: %s line %d\"}\n", #x, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
#define CKD(x) do { CUresult e_ = (x); if (e_ != CUDA_SUCCESS) { const char *s_ = nullptr; cuGetErrorString(e_, &s_); printf("  \"mc_error\": \"%s: %s line %d\",\n", #x, s_ ? s_ : "?", __LINE__); return false; } } while (0)

__device__ __forceinline__ int4 ld16(const void *p) { int4 r; asm volatile("ld.global.cg.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x),"=r"(r.y),"=r"(r.z),"=r"(r.w) : "l"(p)); return r; }
__device__ __forceinline__ void st16(void *p, int4 v) { asm volatile("st.global.cg.v4.s32 [%0], {%1,%2,%3,%4};" :: "l"(p),"r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w) : "memory"); }
__device__ __forceinline__ void mcst16(void *p, int4 v) { asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p),"r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w) : "memory"); }

// MODE 0: dst only; 1: dst + dst2; 2: multimem store to dst
template <int MODE> __global__ void __launch_bounds__(256) copy_k(uint8_t *dst, uint8_t *dst2, const uint8_t *src, size_t n)
{
	const int lane = threadIdx.x & 31;
	size_t warp = (size_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, nw = (size_t)gridDim.x * (blockDim.x / 32);
	for (size_t u = warp; u < n / 4096; u += nw) {
		const uint8_t *s = src + u * 4096;
		int4 r[8];
#pragma unroll
		for (int k = 0; k < 8; k++) r[k] = ld16(s + (lane + 32 * k) * 16);
#pragma unroll
		for (int k = 0; k < 8; k++) {
			const size_t o = u * 4096 + (lane + 32 * k) * 16;
			if (MODE == 2) mcst16(dst + o, r[k]); else st16(dst + o, r[k]);
			if (MODE == 1) st16(dst2 + o, r[k]);
		}
	}
}

__global__ void fill_k(uint64_t *p, size_t nwords, uint64_t seed)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) {
		uint64_t z = (i ^ seed) * 0x9E3779B97F4A7C15ull + 0x9E3779B97F4A7C15ull;
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
		p[i] = z ^ (z >> 31);
	}
}

__global__ void sum_k(const uint64_t *p, size_t nwords, unsigned long long *out)
{
	unsigned long long s = 0;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) s += p[i] * (2 * i + 1);
	atomicAdd(out, s);
}

static unsigned long long digest(int dev, const void *p, size_t n)
{
	CK(cudaSetDevice(dev));
	unsigned long long *d, h = 0;
	CK(cudaMalloc(&d, 8)); CK(cudaMemset(d, 0, 8));
	sum_k<<<592, 256>>>((const uint64_t *)p, n / 8, d);
	CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
	cudaFree(d);
	return h;
}

template <class F> static double best_gbs(size_t bytes, cudaStream_t st, F &&launch, int iters = 4)
{
	cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
	float best = 1e30f;
	for (int it = 0; it < iters; it++) {
		CK(cudaEventRecord(a, st)); launch(); CK(cudaEventRecord(b, st)); CK(cudaEventSynchronize(b));
		float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
	}
	CK(cudaGetLastError());
	cudaEventDestroy(a); cudaEventDestroy(b);
	return bytes / (best * 1e-3) / 1e9;
}

// multicast object over devices 0..ndev-1, `size` bytes, physical memory of its own on every device
static bool multicast_probe(int ndev, size_t size, const uint8_t *src, unsigned long long want)
{
	CKD(cuInit(0));
	int ok = 1;
	for (int d = 0; d < ndev; d++) {
		int v = 0; CUdevice dev; CKD(cuDeviceGet(&dev, d));
		CKD(cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev));
		ok &= v;
	}
	printf("  \"mc_supported\": %d,\n", ok);
	if (!ok) return false;
	CUmulticastObjectProp mp = {};
	mp.numDevices = ndev; mp.size = size; mp.handleTypes = 0; mp.flags = 0;
	size_t gran = 0;
	CKD(cuMulticastGetGranularity(&gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
	size = (size + gran - 1) / gran * gran; mp.size = size;
	CUmemGenericAllocationHandle mc;
	CKD(cuMulticastCreate(&mc, &mp));
	for (int d = 0; d < ndev; d++) { CUdevice dev; CKD(cuDeviceGet(&dev, d)); CKD(cuMulticastAddDevice(mc, dev)); }
	std::vector<CUdeviceptr> uc(ndev), mcva(ndev);
	for (int d = 0; d < ndev; d++) {
		CK(cudaSetDevice(d));
		CUmemAllocationProp ap = {};
		ap.type = CU_MEM_ALLOCATION_TYPE_PINNED; ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE; ap.location.id = d;
		CUmemGenericAllocationHandle h;
		CKD(cuMemCreate(&h, size, &ap, 0));
		CKD(cuMulticastBindMem(mc, 0, h, 0, size, 0));
		CUmemAccessDesc ad = {}; ad.location = ap.location; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
		CKD(cuMemAddressReserve(&uc[d], size, gran, 0, 0)); CKD(cuMemMap(uc[d], size, 0, h, 0)); CKD(cuMemSetAccess(uc[d], size, &ad, 1));
		CKD(cuMemAddressReserve(&mcva[d], size, gran, 0, 0)); CKD(cuMemMap(mcva[d], size, 0, mc, 0)); CKD(cuMemSetAccess(mcva[d], size, &ad, 1));
		CK(cudaMemset((void *)uc[d], 0, size));
		CK(cudaDeviceSynchronize());
	}
	CK(cudaSetDevice(0));
	for (int grid : {148, 296, 592}) {
		double g = best_gbs(size, 0, [&] { copy_k<2><<<grid, 256>>>((uint8_t *)mcva[0], nullptr, src, size); });
		printf("  \"mc_store_r%d_grid%d_gbs\": %.1f,\n", ndev, grid, g);
	}
	CK(cudaDeviceSynchronize());
	int good = 0;
	for (int d = 0; d < ndev; d++) good += digest(d, (void *)uc[d], size) == want;
	printf("  \"mc_replicas_equal_source\": %d,\n", good);
	return true;
}

int main(int argc, char **argv)
{
	int ndev = 0; CK(cudaGetDeviceCount(&ndev));
	size_t n = (argc > 1 ? strtoull(argv[1], nullptr, 0) : 2048) << 20;	// MiB, default 2 GiB
	printf("{\n  \"gpus\": %d, \"bytes\": %zu,\n", ndev, n);
	if (ndev < 2) { printf("  \"note\": \"needs >= 2 GPUs\"\n}\n"); return 0; }
	std::vector<uint8_t *> buf(ndev);
	for (int d = 0; d < ndev; d++) {
		CK(cudaSetDevice(d)); CK(cudaMalloc(&buf[d], n));
		for (int p = 0; p < ndev; p++) if (p != d) { cudaError_t e = cudaDeviceEnablePeerAccess(p, 0); if (e != cudaSuccess) cudaGetLastError(); }
	}
	CK(cudaSetDevice(0));
	uint8_t *src, *local; CK(cudaMalloc(&src, n)); CK(cudaMalloc(&local, n));
	fill_k<<<592, 256>>>((uint64_t *)src, n / 8, 0xC5);
	CK(cudaDeviceSynchronize());
	const unsigned long long want = digest(0, src, n);
	CK(cudaSetDevice(0));

	for (size_t sz : {(size_t)2 << 20, (size_t)64 << 20, n}) {
		double g = best_gbs(sz, 0, [&] { CK(cudaMemcpyPeerAsync(buf[1], 1, src, 0, sz, 0)); }, 6);
		printf("  \"ce_peer_copy_%zuMiB_gbs\": %.1f,\n", sz >> 20, g);
	}
	printf("  \"ce_copy_correct\": %d,\n", (int)(digest(1, buf[1], n) == want));
	CK(cudaSetDevice(0));
	// 2 MiB extents back to back on one stream (the pipelined-replica shape)
	{
		double g = best_gbs(n, 0, [&] { for (size_t o = 0; o < n; o += 2 << 20) CK(cudaMemcpyPeerAsync(buf[1] + o, 1, src + o, 0, 2 << 20, 0)); }, 3);
		printf("  \"ce_peer_copy_2MiB_extents_stream_gbs\": %.1f,\n", g);
	}
	for (int grid : {148, 296, 592, 1184}) {
		CK(cudaMemset(buf[1], 0, n));
		double g = best_gbs(n, 0, [&] { copy_k<0><<<grid, 256>>>(buf[1], nullptr, src, n); });
		printf("  \"sm_store_p2p_grid%d_gbs\": %.1f,\n", grid, g);
	}
	printf("  \"sm_store_correct\": %d,\n", (int)(digest(1, buf[1], n) == want));
	CK(cudaSetDevice(0));
	for (int grid : {296, 592}) {
		double g = best_gbs(n, 0, [&] { copy_k<1><<<grid, 256>>>(local, buf[1], src, n); });
		printf("  \"sm_store_local_plus_p2p_grid%d_gbs\": %.1f,\n", grid, g);
	}
	if (ndev >= 3) {
		// fused fan-out to R-1 = ndev-1 peers by plain stores costs (R-1) x egress: measure R=3 shape with two peers
		double g = best_gbs(n, 0, [&] { copy_k<1><<<592, 256>>>(buf[1], buf[2], src, n); });
		printf("  \"sm_store_two_peers_gbs_payload\": %.1f,\n", g);
	}
	// pull: device 1 reads device 0
	CK(cudaSetDevice(1));
	for (int grid : {296, 592}) {
		double g = best_gbs(n, 0, [&] { copy_k<0><<<grid, 256>>>(buf[1], nullptr, src, n); });
		printf("  \"sm_pull_from_peer_grid%d_gbs\": %.1f,\n", grid, g);
	}
	CK(cudaDeviceSynchronize());
	CK(cudaSetDevice(0));
	// copy engine to every peer at once
	{
		std::vector<cudaStream_t> st(ndev);
		for (int d = 1; d < ndev; d++) CK(cudaStreamCreateWithFlags(&st[d], cudaStreamNonBlocking));
		cudaEvent_t a; CK(cudaEventCreate(&a));
		std::vector<cudaEvent_t> e(ndev);
		for (int d = 1; d < ndev; d++) CK(cudaEventCreate(&e[d]));
		float best = 1e30f;
		for (int it = 0; it < 4; it++) {
			CK(cudaDeviceSynchronize());
			CK(cudaEventRecord(a, st[1]));
			for (int d = 2; d < ndev; d++) CK(cudaStreamWaitEvent(st[d], a, 0));
			for (int d = 1; d < ndev; d++) { CK(cudaMemcpyPeerAsync(buf[d], d, src, 0, n, st[d])); CK(cudaEventRecord(e[d], st[d])); }
			float worst = 0;
			for (int d = 1; d < ndev; d++) { CK(cudaEventSynchronize(e[d])); float ms; CK(cudaEventElapsedTime(&ms, a, e[d])); if (ms > worst) worst = ms; }
			if (worst < best) best = worst;
		}
		printf("  \"ce_fanout_%d_peers_payload_gbs\": %.1f, \"ce_fanout_egress_gbs\": %.1f,\n", ndev - 1, n / (best * 1e-3) / 1e9, (ndev - 1) * (n / (best * 1e-3) / 1e9));
	}
	multicast_probe(ndev, n / 2, src, digest(0, src, n / 2));
	if (ndev > 2) multicast_probe(2, n / 2, src, digest(0, src, n / 2));
	printf("  \"done\": 1\n}\n");
	return 0;
}
