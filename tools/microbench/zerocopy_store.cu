// Bandwidth of SM stores into pinned host memory (the zero-copy output path of gtts_batch_run_host) as a
// function of alignment and store width.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o zerocopy_store zerocopy_store.cu
#include <cstdio>
#include <cuda_runtime.h>

// Each warp writes `rows` rows of 32 floats; row r of warp w starts at (w * rows + r) * 32 + shift floats.
__global__ void store_rows(float* out, long long rowsPerWarp, int shift, int vec)
{
	const int lane = threadIdx.x & 31;
	const long long warp = (long long) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const long long nWarps = (long long) gridDim.x * (blockDim.x >> 5);
	for (long long r = 0; r < rowsPerWarp; ++r) {
		const long long row = r * nWarps + warp;
		if (vec == 1) {
			out[row * 32 + shift + lane] = (float) row;
		} else if (vec == 4) {
			// 8 lanes x float4 = one 128-byte row; 4 rows per instruction
			const long long rr = row * 4 + (lane >> 3);
			reinterpret_cast<float4*>(out + rr * 32 + shift)[lane & 7] = make_float4(1.f, 2.f, 3.f, (float) rr);
		}
	}
}

int main()
{
	const long long bytes = 1ll << 30;
	float* h = nullptr;
	cudaHostAlloc(&h, bytes + 4096, cudaHostAllocMapped);
	float* d = nullptr;
	cudaHostGetDevicePointer(&d, h, 0);
	float* dev = nullptr;
	cudaMalloc(&dev, bytes + 4096);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int grids[] = {148, 148 * 4};
	const int blocks[] = {64, 256, 768};
	for (int target = 0; target < 2; ++target)
	for (int gi = 0; gi < 2; ++gi)
	for (int bi = 0; bi < 3; ++bi)
	for (int shift = 0; shift <= 13; shift += 13)
	for (int vec = 1; vec <= 4; vec += 3) {
		if (vec == 4 && shift % 4) continue;
		const long long nWarps = (long long) grids[gi] * (blocks[bi] / 32);
		const long long rowsTotal = bytes / 128;
		const long long rowsPerWarp = rowsTotal / nWarps / (vec == 4 ? 4 : 1);
		float ms = 1e9f;
		for (int rep = 0; rep < 3; ++rep) {
			cudaEventRecord(e0);
			store_rows<<<grids[gi], blocks[bi]>>>(target ? dev : d, rowsPerWarp, shift, vec);
			cudaEventRecord(e1);
			cudaEventSynchronize(e1);
			float t; cudaEventElapsedTime(&t, e0, e1);
			if (t < ms) ms = t;
		}
		printf("%s grid %4d block %4d shift %2d vec %d: %7.2f ms  %6.1f GB/s\n", target ? "device" : "host  ",
		       grids[gi], blocks[bi], shift, vec, ms, (double) bytes / 1e6 / ms);
	}
	cudaError_t e = cudaGetLastError();
	printf("status: %s\n", cudaGetErrorString(e));
	return 0;
}
