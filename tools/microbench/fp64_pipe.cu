// Micro-benchmark of the pipes the tube path is bounded by:
//   * FP64 FMA peak (MEASURED_PEAKS.json has only HBM and bf16) -> the roofline denominator of bench.py
//   * dependent-chain latencies (DFMA, DADD, DMUL, SHFL, LDS) -> the serial tube recurrence budget
//   * cost of partially active warps on the FP64 pipe
//   * FP64 division / exp2 throughput
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_pipe fp64_pipe.cu
// Prints one JSON object.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

template<int ILP>
__global__ void k_dfma_tput(double* out, int iters, double a, double b)
{
	double x[ILP];
#pragma unroll
	for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-9 + i;
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
	}
	double s = 0;
#pragma unroll
	for (int i = 0; i < ILP; i++) s += x[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<int ILP>
__global__ void k_ffma_tput(float* out, int iters, float a, float b)
{
	float x[ILP];
#pragma unroll
	for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-9f + i;
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int i = 0; i < ILP; i++) x[i] = fmaf(x[i], a, b);
	}
	float s = 0;
#pragma unroll
	for (int i = 0; i < ILP; i++) s += x[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 0 DFMA chain, 1 DADD chain, 2 DMUL chain, 3 SHFL(64-bit up) chain, 4 LDS chain, 5 DADD->DSETP->SEL chain,
// 6 division chain, 7 exp2 chain, 8 FADD chain
__global__ void k_latency(double* out, long long* cycles, int iters, int mode, double a, double b, unsigned mask)
{
	__shared__ double sm[64];
	const int lane = threadIdx.x & 31;
	sm[lane] = 0.0; sm[lane + 32] = 0.0;
	__syncthreads();
	if (!((mask >> lane) & 1)) return;
	double x = lane * 1e-3 + 1.0;
	float xf = lane * 1e-3f + 1.0f;
	int idx = lane;
	long long t0 = clock64();
	if (mode == 0) { for (int i = 0; i < iters; i++) { x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); } }
	else if (mode == 1) { for (int i = 0; i < iters; i++) { x = __dadd_rn(x, b); x = __dadd_rn(x, b); x = __dadd_rn(x, b); x = __dadd_rn(x, b); } }
	else if (mode == 2) { for (int i = 0; i < iters; i++) { x = __dmul_rn(x, a); x = __dmul_rn(x, a); x = __dmul_rn(x, a); x = __dmul_rn(x, a); } }
	else if (mode == 3) { for (int i = 0; i < iters; i++) { x = __shfl_up_sync(mask, x, 1, 16); x = __shfl_up_sync(mask, x, 1, 16); x = __shfl_up_sync(mask, x, 1, 16); x = __shfl_up_sync(mask, x, 1, 16); } }
	else if (mode == 4) { for (int i = 0; i < iters; i++) { idx = (int) sm[idx]; idx = (int) sm[idx + 1]; idx = (int) sm[idx]; idx = (int) sm[idx + 1]; } x = idx; }
	else if (mode == 5) { for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int r = 0; r < 4; r++) { double s = __dadd_rn(x, b); double t = __dadd_rn(s, -512.0); x = (s > 511.0) ? t : s; } } }
	else if (mode == 6) { for (int i = 0; i < iters; i++) { x = a / x; x = a / x; x = a / x; x = a / x; } }
	else if (mode == 7) { for (int i = 0; i < iters; i++) { x = exp2(x * 1e-3); x = exp2(x * 1e-3); x = exp2(x * 1e-3); x = exp2(x * 1e-3); } }
	else if (mode == 8) { for (int i = 0; i < iters; i++) { xf = __fadd_rn(xf, (float) b); xf = __fadd_rn(xf, (float) b); xf = __fadd_rn(xf, (float) b); xf = __fadd_rn(xf, (float) b); } x = xf; }
	long long t1 = clock64();
	out[threadIdx.x] = x;
	if (lane == 0 || ((mask & 1) == 0 && lane == 16)) cycles[0] = t1 - t0;
}

// Throughput of a DP instruction mix with a given active-lane mask, many warps per SM.
__global__ void k_mask_tput(double* out, int iters, double a, double b, unsigned mask)
{
	const int lane = threadIdx.x & 31;
	double x[8];
#pragma unroll
	for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-9 + i;
	if ((mask >> lane) & 1) {
		for (int it = 0; it < iters; it++) {
#pragma unroll
			for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
		}
	}
	double s = 0;
#pragma unroll
	for (int i = 0; i < 8; i++) s += x[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// throughput of division / exp2 / pow with many warps
__global__ void k_special_tput(double* out, int iters, int mode, double a)
{
	double x[4];
#pragma unroll
	for (int i = 0; i < 4; i++) x[i] = 1.0 + threadIdx.x * 1e-6 + i * 0.25;
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int i = 0; i < 4; i++) {
			if (mode == 0) x[i] = a / x[i];
			else if (mode == 1) x[i] = exp2(x[i] * 0.01);
			else if (mode == 2) x[i] = pow(10.0, x[i] * -0.01) + 1.0;
			else if (mode == 3) x[i] = tan(x[i] * 0.01) + 1.0;
			else if (mode == 4) x[i] = cos(x[i] * 0.01) + 1.0;
			else x[i] = exp10(x[i] * -0.01) + 1.0;
		}
	}
	double s = 0;
#pragma unroll
	for (int i = 0; i < 4; i++) s += x[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float timeit(void (*launch)(void*), void* arg, int reps)
{
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	for (int i = 0; i < 3; i++) launch(arg);
	CK(cudaDeviceSynchronize());
	CK(cudaEventRecord(e0));
	for (int i = 0; i < reps; i++) launch(arg);
	CK(cudaEventRecord(e1));
	CK(cudaEventSynchronize(e1));
	float ms;
	CK(cudaEventElapsedTime(&ms, e0, e1));
	return ms / reps;
}

struct Args { double* d; float* f; int blocks, threads, iters; unsigned mask; int mode; };
static void l_dfma8(void* p) { Args* a = (Args*) p; k_dfma_tput<8><<<a->blocks, a->threads>>>(a->d, a->iters, 1.0000001, 1e-9); }
static void l_dfma2(void* p) { Args* a = (Args*) p; k_dfma_tput<2><<<a->blocks, a->threads>>>(a->d, a->iters, 1.0000001, 1e-9); }
static void l_ffma8(void* p) { Args* a = (Args*) p; k_ffma_tput<8><<<a->blocks, a->threads>>>(a->f, a->iters, 1.0000001f, 1e-9f); }
static void l_mask(void* p) { Args* a = (Args*) p; k_mask_tput<<<a->blocks, a->threads>>>(a->d, a->iters, 1.0000001, 1e-9, a->mask); }
static void l_special(void* p) { Args* a = (Args*) p; k_special_tput<<<a->blocks, a->threads>>>(a->d, a->iters, a->mode, 1.5); }

int main()
{
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, 0));
	const int sms = prop.multiProcessorCount;
	int clk_khz = 0;
	CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
	double* d; float* f; long long* cyc;
	CK(cudaMalloc(&d, sizeof(double) * sms * 16 * 1024));
	CK(cudaMalloc(&f, sizeof(float) * sms * 16 * 1024));
	CK(cudaMalloc(&cyc, sizeof(long long) * 4));
	printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d", prop.name, sms, clk_khz);

	Args a{d, f, sms * 8, 512, 4096, 0xffffffffu, 0};
	{
		float ms = timeit(l_dfma8, &a, 10);
		double flops = 2.0 * 8 * a.iters * (double) a.blocks * a.threads;
		printf(", \"fp64_fma_tflops\": %.3f", flops / ms * 1e-9);
		ms = timeit(l_dfma2, &a, 10);
		flops = 2.0 * 2 * a.iters * (double) a.blocks * a.threads;
		printf(", \"fp64_fma_tflops_ilp2\": %.3f", flops / ms * 1e-9);
		ms = timeit(l_ffma8, &a, 10);
		flops = 2.0 * 8 * a.iters * (double) a.blocks * a.threads;
		printf(", \"fp32_fma_tflops\": %.3f", flops / ms * 1e-9);
	}
	// sustained (several seconds) FP64 figure, for a kernel timed inside a long step
	{
		Args s = a; s.iters = 65536;
		float ms = timeit(l_dfma8, &s, 40);
		double flops = 2.0 * 8 * s.iters * (double) s.blocks * s.threads;
		printf(", \"fp64_fma_tflops_sustained\": %.3f, \"sustained_ms_per_launch\": %.2f", flops / ms * 1e-9, ms);
	}
	// active-mask sweep: warp-instructions per second when only some lanes are active
	{
		unsigned masks[] = {0xffffffffu, 0x0000ffffu, 0x000000ffu, 0x00ff00ffu, 0x55555555u, 0x00000001u};
		printf(", \"mask_sweep_gwarpinst_per_s\": {");
		for (int i = 0; i < 6; i++) {
			a.mask = masks[i];
			float ms = timeit(l_mask, &a, 10);
			double winst = 8.0 * a.iters * (double) a.blocks * a.threads / 32;
			printf("%s\"%08x\": %.2f", i ? ", " : "", masks[i], winst / ms * 1e-6);
		}
		printf("}");
	}
	// specials throughput (G ops/s, all lanes active)
	{
		const char* names[] = {"div", "exp2", "pow10", "tan", "cos", "exp10"};
		Args s = a; s.iters = 256;
		printf(", \"special_gops\": {");
		for (int m = 0; m < 6; m++) {
			s.mode = m;
			float ms = timeit(l_special, &s, 5);
			double ops = 4.0 * s.iters * (double) s.blocks * s.threads;
			printf("%s\"%s\": %.2f", m ? ", " : "", names[m], ops / ms * 1e-6);
		}
		printf("}");
	}
	// latencies (cycles per dependent op), single warp
	{
		const char* names[] = {"dfma", "dadd", "dmul", "shfl64", "lds_cvt", "dadd_setp_sel", "ddiv", "exp2", "fadd"};
		printf(", \"latency_cycles\": {");
		for (int m = 0; m < 9; m++) {
			const int iters = 2048;
			k_latency<<<1, 32>>>(d, cyc, iters, m, 1.0000001, 1e-9, 0xffffffffu);
			CK(cudaDeviceSynchronize());
			k_latency<<<1, 32>>>(d, cyc, iters, m, 1.0000001, 1e-9, 0xffffffffu);
			CK(cudaDeviceSynchronize());
			long long c;
			CK(cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost));
			printf("%s\"%s\": %.2f", m ? ", " : "", names[m], (double) c / (4.0 * iters));
		}
		printf("}");
		// same DFMA chain with half / quarter masks (does a half-empty warp issue faster?)
		printf(", \"latency_dfma_masked\": {");
		unsigned masks[] = {0x0000ffffu, 0x000000ffu};
		for (int i = 0; i < 2; i++) {
			k_latency<<<1, 32>>>(d, cyc, 2048, 0, 1.0000001, 1e-9, masks[i]);
			CK(cudaDeviceSynchronize());
			long long c;
			CK(cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost));
			printf("%s\"%08x\": %.2f", i ? ", " : "", masks[i], (double) c / (4.0 * 2048));
		}
		printf("}");
	}
	printf("}\n");
	return 0;
}
