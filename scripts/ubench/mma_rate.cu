// micro-benchmark: issue rate of the legacy warp-level tensor-core instructions on sm_100a (per SM, all 4 schedulers busy)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(int iters, int* out) {
	int c[4][4] = {};
	float f[4][4] = {};
	unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int u = 0; u < 4; u++) {
			if (MODE == 0)
				asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
				             : "+r"(c[u][0]), "+r"(c[u][1]), "+r"(c[u][2]), "+r"(c[u][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
			else if (MODE == 1)
				asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
				             : "+f"(f[u][0]), "+f"(f[u][1]), "+f"(f[u][2]), "+f"(f[u][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
			else if (MODE == 2)
				asm volatile("mma.sync.aligned.m16n8k32.row.col.f32.e4m3.e4m3.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
				             : "+f"(f[u][0]), "+f"(f[u][1]), "+f"(f[u][2]), "+f"(f[u][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
			else {
				asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(c[u][0]) : "r"(a0), "r"(b0));
				asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(c[u][1]) : "r"(a1), "r"(b1));
				asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(c[u][2]) : "r"(a2), "r"(b0));
				asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(c[u][3]) : "r"(a3), "r"(b1));
			}
		}
	}
	int s = 0;
	for (int u = 0; u < 4; u++) for (int j = 0; j < 4; j++) s += c[u][j] + (int) f[u][j];
	if (s == 12345) *out = s;
}
template <int MODE>
void run(const char* name, double macs_per_instr, int per_iter) {
	int* d; cudaMalloc(&d, 4);
	const int iters = 20000, threads = 512, blocks = 148;
	k<MODE><<<blocks, threads>>>(100, d);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	k<MODE><<<blocks, threads>>>(iters, d);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	const double instr_per_sm = (double) iters * per_iter * (threads / 32);
	const double cyc = ms * 1e-3 * 1.965e9;
	printf("%-28s %8.3f ms  %6.2f cycles per warp-instruction per SM  (%7.0f MAC/clk/SM)  err=%s\n", name, ms, cyc / instr_per_sm,
	       macs_per_instr * instr_per_sm / cyc, cudaGetErrorString(cudaGetLastError()));
}
int main() {
	run<0>("IMMA m16n8k32 u8.u8.s32", 16 * 8 * 32, 4);
	run<1>("HMMA m16n8k16 f16.f16.f32", 16 * 8 * 16, 4);
	run<2>("QMMA m16n8k32 e4m3.e4m3.f32", 16 * 8 * 32, 4);
	run<3>("IDP.4A (dp4a)", 32 * 4, 16);
	return 0;
}
