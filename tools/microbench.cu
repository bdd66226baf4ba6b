// tools/microbench.cu -- issue-rate probe for the instructions the fill kernel is made of (B200, sm_100a).
// For each op: every warp runs ILP independent chains of the op in a long unrolled loop; cycles are read with
// clock64() around the loop; reported figure = warp-instructions per clock per SM with WARPS resident warps.
// Output: one JSON object on stdout (kept under profiles/ as dpx_peak.json).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define ILP 8

template <int OP>
__device__ __forceinline__ uint32_t apply(uint32_t a, uint32_t b, uint32_t c, int lane)
{
	if (OP == 0) return __viaddmax_s16x2(a, b, c);
	if (OP == 1) return __viaddmax_s16x2_relu(a, b, c);
	if (OP == 2) return __vmaxs2(a, b);
	if (OP == 3) return __vimax3_s16x2(a, b, c);
	if (OP == 4) return __vadd2(a, b);
	if (OP == 5) return a * b + c;                       // IMAD
	if (OP == 6) return __shfl_up_sync(0xffffffffu, a, 1);
	if (OP == 7) return (a & b) ^ c;                     // LOP3
	if (OP == 8) return a + b + c;                       // IADD3
	if (OP == 9) { uint32_t d; asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }   // HMNMX2
	return a;
}

template <int OP>
__global__ void probe(uint32_t* out, long long* cyc, uint32_t seed)
{
	uint32_t v[ILP];
	const int lane = threadIdx.x & 31;
#pragma unroll
	for (int i = 0; i < ILP; ++i) v[i] = seed * (i + 1) + threadIdx.x;
	uint32_t b = seed | 1, c = seed >> 3;
	__syncthreads();
	long long t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < ITERS; ++it) {
#pragma unroll
		for (int u = 0; u < 4; ++u)
#pragma unroll
			for (int i = 0; i < ILP; ++i) v[i] = apply<OP>(v[i], OP == 5 ? b : v[(i + 1) % ILP], OP == 5 ? c : v[(i + 3) % ILP], lane);
	}
	long long t1 = clock64();
	uint32_t acc = 0;
#pragma unroll
	for (int i = 0; i < ILP; ++i) acc ^= v[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// mixed: per iteration A DPX ops and B IMADs on independent chains -> does the second pipe add throughput?
template <int NA, int NB>
__global__ void probe_mix(uint32_t* out, long long* cyc, uint32_t seed)
{
	uint32_t va[8], vb[8];
#pragma unroll
	for (int i = 0; i < 8; ++i) { va[i] = seed * (i + 1) + threadIdx.x; vb[i] = seed * (i + 9) + threadIdx.x; }
	uint32_t b = seed | 1, c = seed >> 3;
	__syncthreads();
	long long t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < ITERS; ++it) {
#pragma unroll
		for (int u = 0; u < 4; ++u) {
#pragma unroll
			for (int i = 0; i < NA; ++i) va[i] = __viaddmax_s16x2(va[i], va[(i + 1) % NA], va[(i + 3) % NA]);
#pragma unroll
			for (int i = 0; i < NB; ++i) vb[i] = vb[i] * b + c;
		}
	}
	long long t1 = clock64();
	uint32_t acc = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i) acc ^= va[i] ^ vb[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// mixed: NA DPX ops and NB packed-half maxima (HMNMX2, FMA-side pipe) per iteration on independent chains
template <int NA, int NB>
__global__ void probe_mix_h(uint32_t* out, long long* cyc, uint32_t seed)
{
	uint32_t va[8], vb[8];
#pragma unroll
	for (int i = 0; i < 8; ++i) { va[i] = seed * (i + 1) + threadIdx.x; vb[i] = (seed * (i + 9) + threadIdx.x) & 0x3fff3fffu; }
	__syncthreads();
	long long t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < ITERS; ++it) {
#pragma unroll
		for (int u = 0; u < 4; ++u) {
#pragma unroll
			for (int i = 0; i < NA; ++i) va[i] = __viaddmax_s16x2(va[i], va[(i + 1) % NA], va[(i + 3) % NA]);
#pragma unroll
			for (int i = 0; i < NB; ++i) { uint32_t d; asm volatile("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(vb[i]), "r"(vb[(i + 1) % NB])); vb[i] = d; }
		}
	}
	long long t1 = clock64();
	uint32_t acc = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i) acc ^= va[i] ^ vb[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <class K>
static double run(K kern, int threads, int ops_per_iter, int sms)
{
	uint32_t* out; long long* cyc;
	cudaMalloc(&out, sizeof(uint32_t) * threads * sms);
	cudaMalloc(&cyc, sizeof(long long) * sms);
	kern<<<sms, threads>>>(out, cyc, 12345u);
	kern<<<sms, threads>>>(out, cyc, 12345u);
	cudaDeviceSynchronize();
	long long* h = new long long[sms];
	cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
	double avg = 0;
	for (int i = 0; i < sms; ++i) avg += (double)h[i];
	avg /= sms;
	delete[] h;
	cudaFree(out); cudaFree(cyc);
	const double warp_instr = (double)(threads / 32) * ITERS * ops_per_iter;
	return warp_instr / avg;          // warp-instructions per clock per SM
}

int main()
{
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	const int sms = p.multiProcessorCount;
	int clk_khz = 0;
	cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
	printf("{\"device\": \"%s\", \"sms\": %d, \"max_sm_khz\": %d", p.name, sms, clk_khz);
	const char* names[] = {"VIADDMNMX.S16x2", "VIADDMNMX.S16x2.RELU", "VIMNMX.S16x2", "VIMNMX3.S16x2", "VIADD.16x2", "IMAD", "SHFL.UP", "LOP3", "IADD3"};
	for (int threads = 256; threads <= 1024; threads *= 2) {
		double r[9];
		r[0] = run(probe<0>, threads, 4 * ILP, sms);
		r[1] = run(probe<1>, threads, 4 * ILP, sms);
		r[2] = run(probe<2>, threads, 4 * ILP, sms);
		r[3] = run(probe<3>, threads, 4 * ILP, sms);
		r[4] = run(probe<4>, threads, 4 * ILP, sms);
		r[5] = run(probe<5>, threads, 4 * ILP, sms);
		r[6] = run(probe<6>, threads, 4 * ILP, sms);
		r[7] = run(probe<7>, threads, 4 * ILP, sms);
		r[8] = run(probe<8>, threads, 4 * ILP, sms);
		printf(", \"warp_instr_per_clk_per_sm_%dthr\": {", threads);
		for (int i = 0; i < 9; ++i) printf("%s\"%s\": %.3f", i ? ", " : "", names[i], r[i]);
		printf("}");
		if (threads == 1024) {
			const double m44 = run(probe_mix<4, 4>, threads, 4 * 8, sms);
			const double m62 = run(probe_mix<6, 2>, threads, 4 * 8, sms);
			const double m80 = run(probe_mix<8, 0>, threads, 4 * 8, sms);
			const double h80 = run(probe_mix_h<8, 0>, threads, 4 * 8, sms);
			const double h62 = run(probe_mix_h<6, 2>, threads, 4 * 8, sms);
			const double h44 = run(probe_mix_h<4, 4>, threads, 4 * 8, sms);
			const double h08 = run(probe<9>, threads, 4 * ILP, sms);
			printf(", \"mix_dpx_hmnmx2_1024thr\": {\"8+0\": %.3f, \"6+2\": %.3f, \"4+4\": %.3f, \"HMNMX2 alone\": %.3f}", h80, h62, h44, h08);
			printf(", \"mix_dpx_imad_1024thr\": {\"4+4\": %.3f, \"6+2\": %.3f, \"8+0\": %.3f}", m44, m62, m80);
			// fill kernel: 5.5 DPX ops per lane per two cells -> cells/clk/SM = rate * 32 lanes * 2 / 5.5
			const double cells_per_clk_sm = r[0] * 32.0 * 2.0 / 5.5;
			printf(", \"gcups_peak_5p5_ops_per_cellpair\": %.1f", cells_per_clk_sm * sms * (clk_khz * 1e3) / 1e9);
		}
	}
	printf("}\n");
	return 0;
}
