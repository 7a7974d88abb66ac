// What bounds one sample of the tube loop on a single warp: dependent FP64 chain + 64-bit shuffles.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tube_chain tube_chain.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src, 32); }

// K dependent DFMAs, then NS shuffles of the result (all feeding the next iteration), plus E independent DFMAs
template<int K, int NS, int E>
__global__ void chain(double* out, long long* cycles, int iters, double k, double c)
{
	const int lane = threadIdx.x & 31;
	double x = 1.0 + lane * 1e-3, y = 0.5, z = 0.25;
	double e[E > 0 ? E : 1];
	for (int i = 0; i < E; ++i) e[i] = 0.1 * i + lane;
	const int s1 = (lane + 31) & 31, s2 = (lane + 1) & 31, s3 = lane ^ 4;
	const long long t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < iters; ++it) {
		double v = x + y + z;
#pragma unroll
		for (int i = 0; i < K; ++i) v = fma(v, k, c);
#pragma unroll
		for (int i = 0; i < E; ++i) e[i] = fma(e[i], k, c);
		if (NS >= 1) x = shfl_d(v, s1); else x = v;
		if (NS >= 2) y = shfl_d(v * k, s2);
		if (NS >= 3) z = shfl_d(v * c, s3);
	}
	const long long t1 = clock64();
	double acc = x + y + z;
	for (int i = 0; i < E; ++i) acc += e[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
	if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template<int K, int NS, int E>
void run(const char* name, double* out, long long* cyc)
{
	const int iters = 20000;
	chain<K, NS, E><<<1, 32>>>(out, cyc, iters, 0.999, 1e-3);
	cudaDeviceSynchronize();
	long long h;
	cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
	printf("%-40s K=%2d shuffles=%d extra=%2d: %7.1f cycles per iteration\n", name, K, NS, E, (double) h / iters);
}

int main()
{
	double* out; long long* cyc;
	cudaMalloc(&out, 1024 * sizeof(double)); cudaMalloc(&cyc, 64 * sizeof(long long));
	run<4, 0, 0>("chain only", out, cyc);
	run<4, 1, 0>("chain + 1 shuffle", out, cyc);
	run<4, 2, 0>("chain + 2 shuffles", out, cyc);
	run<4, 3, 0>("chain + 3 shuffles", out, cyc);
	run<8, 3, 0>("longer chain + 3 shuffles", out, cyc);
	run<4, 3, 8>("chain + 3 shuffles + 8 independent", out, cyc);
	run<4, 3, 16>("chain + 3 shuffles + 16 independent", out, cyc);
	run<0, 3, 0>("3 shuffles only", out, cyc);
	run<0, 1, 0>("1 shuffle only", out, cyc);
	run<0, 0, 16>("16 independent DFMA", out, cyc);
	run<0, 0, 32>("32 independent DFMA", out, cyc);
	printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
