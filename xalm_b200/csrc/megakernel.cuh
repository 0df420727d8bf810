// megakernel.cuh — every layer of one decode token in ONE persistent kernel.
//
// Why (profiles/r1_timeline_multikernel.md): with one kernel per fused op (5 per layer, 160 per token) the in-kernel
// timeline shows ~44 us of streaming per Mistral-7B layer and ~30 us of launch / dependency / prologue gaps between
// kernels — the 3.5-5 us each boundary costs is as long as the small matvecs themselves, and HBM idles through it.
// Here a token's layers are a list of PHASES executed by the same resident CTAs (one per SM):
//
//     [norm+QKV+rope+KV] -> [attention] -> [Wo+residual] -> [norm+W1|W3+GLU] -> [W2+residual] -> next layer ...
//
//   * phases are separated by a software grid barrier among the consumer warps (atomic counter in global memory);
//   * the PRODUCER warp never waits for a barrier: weights are immutable, so it walks the token's whole tile list and
//     keeps the shared-memory ring (~140 KB per SM, ~20 MB chip-wide) full with the NEXT phase's weights while the
//     consumers are still synchronising / staging activations — HBM keeps streaming through every phase boundary;
//   * activations written by other SMs inside the kernel are read with ld.global.cg (L2), never through L1.
//
// Mapping of one matvec phase.  A stage is RCS rows x one 16-byte piece per lane per K-slice warp: KW K-slice warps x RW
// row groups = 16 consumer warps, every warp owns 2 rows of the stage (a RoPE pair / a W1,W3 pair) and ONE piece per row,
// so for n = 4096 (QKV, Wo, W1|W3: 3/4 of the bytes) the activation slice of a lane never changes and lives in
// registers for the whole phase.  K-slice partial sums are combined per tile through shared memory in a fixed order.
//     8-bit / fp16 / fp32 formats: KW = 8, RW = 2, RCS = 4      4- and 5-bit formats: KW = 4, RW = 4, RCS = 8
#pragma once
#include "attention.cuh"
#include "matvec_tma.cuh"

namespace xalm {

constexpr int MK_CW = 16;                 // consumer warps
constexpr int MK_THREADS = (MK_CW + 1) * 32;

template <int TYPE>
struct MkCfg {
	static constexpr int PPU = UFmt<TYPE>::PPU;
	static constexpr int KW = PPU == 8 ? 4 : 8;
	static constexpr int RW = MK_CW / KW;
	static constexpr int RCS = 2 * RW;               // rows per stage (= per tile)
	static constexpr int PIECES = KW * 32;           // pieces per row per stage
	static constexpr int U = PIECES / PPU > 0 ? PIECES / PPU : 1; // units per stage (f32: 4, f16: 8, 8-bit: 16, 4-bit: 16)
};

enum { MK_MATVEC = 0, MK_ATTN = 1 };

struct MkPhase {
	int kind;
	int n_tiles;   // matvec: virtual rows / RCS
	int kranges;   // matvec: stages per tile
	int pad;
	MatvecArgs a;
	AttnArgs at;
	int G, HD;
};

// virtual row r of tile -> physical weight row.  GLU tiles interleave W1/W3 rows so each 2-row group holds (W1[o], W3[o]).
__device__ __forceinline__ int mk_phys_row(const MatvecArgs& a, int row0, int r) {
	if (a.epi == EPI_GLU) {
		const int o = (row0 >> 1) + (r >> 1);
		return (r & 1) ? a.glu_off + o : o;
	}
	return row0 + r;
}

__device__ __forceinline__ float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void mk_bar_sync_all() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void mk_bar_sync_group(int g, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(2 + g), "r"(nthreads) : "memory"); }

// grid barrier among the consumer warps of all CTAs: monotonically increasing arrival counter
__device__ __forceinline__ void mk_grid_barrier(unsigned int* counter, unsigned int target) {
	mk_bar_sync_all();
	if (threadIdx.x == 0) {
		__threadfence();
		atomicAdd(counter, 1u);
		unsigned int v;
		do {
			asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
		} while (v < target);
		__threadfence();
	}
	mk_bar_sync_all();
}

// ---- attention phase: the 16 consumer warps of a CTA process (kv head, split) items; same math as attn_decode_kernel ----
template <int HD, int G>
__device__ __noinline__ void mk_attention(const AttnArgs& a, float* scratch) {
	constexpr int NW = MK_CW;
	constexpr int LPR = HD / 8, RPW = 32 / LPR, TB = 4, NGRP = NW * RPW;
	float* s_m = scratch;                       // [NGRP][G]
	float* s_l = s_m + NGRP * G;                // [NGRP][G]
	float* s_scale = s_l + NGRP * G;            // [NGRP][G]
	float* s_acc = s_scale + NGRP * G;          // [NW][G][HD]
	__shared__ int s_last;
	const int kv_len = a.kv_len_fixed >= 0 ? a.kv_len_fixed : a.step->kv_len;
	const int slen = attn_split_len(kv_len, a.n_splits, a.min_split);
	const int n_active = (kv_len + slen - 1) / slen;
	const int n_items = a.n_kv_heads * n_active;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int sub = lane / LPR, li = lane % LPR;
	const int kv_stride = a.n_kv_heads * HD;
	const float inv_sqrt = 1.0f / sqrtf((float) HD);
	for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
		const int kvh = item / n_active, split = item % n_active;
		const int t0 = split * slen, t1 = min(kv_len, t0 + slen);
		float qf[G][8];
#pragma unroll
		for (int g = 0; g < G; g++) {
			const float* qp = a.q + (size_t) (kvh * G + g) * HD + li * 8;
			const float4 u = ld_cg4(qp), v = ld_cg4(qp + 4);
			qf[g][0] = u.x; qf[g][1] = u.y; qf[g][2] = u.z; qf[g][3] = u.w;
			qf[g][4] = v.x; qf[g][5] = v.y; qf[g][6] = v.z; qf[g][7] = v.w;
		}
		float m[G], l[G], acc[G][8];
#pragma unroll
		for (int g = 0; g < G; g++) {
			m[g] = -CUDART_INF_F; l[g] = 0.f;
#pragma unroll
			for (int i = 0; i < 8; i++) acc[g][i] = 0.f;
		}
		const __half* kbase = a.k_cache + (size_t) kvh * HD + li * 8;
		const __half* vbase = a.v_cache + (size_t) kvh * HD + li * 8;
		for (int tb = t0 + warp * RPW * TB; tb < t1; tb += NW * RPW * TB) {
			uint4 kq[TB], vq[TB];
			bool ok[TB];
#pragma unroll
			for (int j = 0; j < TB; j++) {
				const int t = tb + j * RPW + sub;
				ok[j] = t < t1;
				const int tc = ok[j] ? t : t0;
				kq[j] = __ldcg(reinterpret_cast<const uint4*>(kbase + (size_t) tc * kv_stride));
				vq[j] = __ldcg(reinterpret_cast<const uint4*>(vbase + (size_t) tc * kv_stride));
			}
			float s[TB][G];
#pragma unroll
			for (int j = 0; j < TB; j++) {
				const __half2* kh = reinterpret_cast<const __half2*>(&kq[j]);
				float kf[8];
#pragma unroll
				for (int i = 0; i < 4; i++) {
					const float2 f = __half22float2(kh[i]);
					kf[2 * i] = f.x; kf[2 * i + 1] = f.y;
				}
#pragma unroll
				for (int g = 0; g < G; g++) {
					float p = 0.f;
#pragma unroll
					for (int i = 0; i < 8; i++) p += qf[g][i] * kf[i];
#pragma unroll
					for (int o = LPR / 2; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
					s[j][g] = ok[j] ? p * inv_sqrt : -CUDART_INF_F;
				}
			}
#pragma unroll
			for (int g = 0; g < G; g++) {
				float mn = m[g];
#pragma unroll
				for (int j = 0; j < TB; j++) mn = fmaxf(mn, s[j][g]);
				if (mn == -CUDART_INF_F) continue;
				const float corr = expf(m[g] - mn);
				l[g] *= corr;
#pragma unroll
				for (int i = 0; i < 8; i++) acc[g][i] *= corr;
				m[g] = mn;
#pragma unroll
				for (int j = 0; j < TB; j++) {
					const float p = expf(s[j][g] - mn);
					l[g] += p;
					const __half2* vh = reinterpret_cast<const __half2*>(&vq[j]);
#pragma unroll
					for (int i = 0; i < 4; i++) {
						const float2 f = __half22float2(vh[i]);
						acc[g][2 * i] += p * f.x;
						acc[g][2 * i + 1] += p * f.y;
					}
				}
			}
		}
		// ---- merge lane groups and warps ----
		const int grp = warp * RPW + sub;
		if (li == 0) {
#pragma unroll
			for (int g = 0; g < G; g++) { s_m[grp * G + g] = m[g]; s_l[grp * G + g] = l[g]; }
		}
		mk_bar_sync_all();
#pragma unroll
		for (int g = 0; g < G; g++) {
			float M = -CUDART_INF_F;
			for (int i = 0; i < NGRP; i++) M = fmaxf(M, s_m[i * G + g]);
			const float sc = m[g] == -CUDART_INF_F ? 0.f : expf(m[g] - M);
			if (li == 0) s_scale[grp * G + g] = sc;
#pragma unroll
			for (int i = 0; i < 8; i++) {
				float v = acc[g][i] * sc;
#pragma unroll
				for (int o = LPR; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
				acc[g][i] = v;
			}
			if (sub == 0) {
#pragma unroll
				for (int i = 0; i < 8; i++) s_acc[(warp * G + g) * HD + li * 8 + i] = acc[g][i];
			}
		}
		mk_bar_sync_all();
		const bool single = n_active == 1; // one split: write the normalised output directly, no partial round trip
		float* pacc = a.part_acc + ((size_t) kvh * a.n_splits + split) * G * HD;
		float* pml = a.part_ml + ((size_t) kvh * a.n_splits + split) * G * 2;
		for (int i = threadIdx.x; i < G * HD; i += NW * 32) {
			const int g = i / HD, dpos = i % HD;
			float v = 0.f;
#pragma unroll
			for (int w = 0; w < NW; w++) v += s_acc[(w * G + g) * HD + dpos];
			if (single) {
				float L = 0.f;
				for (int k = 0; k < NGRP; k++) L += s_l[k * G + g] * s_scale[k * G + g];
				a.out[(size_t) kvh * G * HD + i] = v / L;
			} else {
				pacc[i] = v;
			}
		}
		if (!single && threadIdx.x < G) {
			const int g = threadIdx.x;
			float L = 0.f, M = -CUDART_INF_F;
			for (int k = 0; k < NGRP; k++) { L += s_l[k * G + g] * s_scale[k * G + g]; M = fmaxf(M, s_m[k * G + g]); }
			pml[2 * g] = M;
			pml[2 * g + 1] = L;
		}
		if (!single) {
			__threadfence();
			mk_bar_sync_all();
			if (threadIdx.x == 0) {
				const unsigned int ticket = atomicAdd(&a.tickets[kvh], 1u);
				s_last = ticket == (unsigned int) (n_active - 1);
				if (s_last) a.tickets[kvh] = 0;
			}
			mk_bar_sync_all();
			if (s_last) {
				__threadfence();
				const float* bacc = a.part_acc + (size_t) kvh * a.n_splits * G * HD;
				const float* bml = a.part_ml + (size_t) kvh * a.n_splits * G * 2;
				for (int i = threadIdx.x; i < G * HD; i += NW * 32) {
					const int g = i / HD;
					float mm = -CUDART_INF_F;
					for (int sidx = 0; sidx < n_active; sidx++) mm = fmaxf(mm, __ldcg(bml + ((size_t) sidx * G + g) * 2));
					float num = 0.f, den = 0.f;
					for (int sidx = 0; sidx < n_active; sidx++) {
						const float ms = __ldcg(bml + ((size_t) sidx * G + g) * 2), ls = __ldcg(bml + ((size_t) sidx * G + g) * 2 + 1);
						const float sc = expf(ms - mm);
						num += sc * __ldcg(bacc + (size_t) sidx * G * HD + i);
						den += sc * ls;
					}
					a.out[(size_t) kvh * G * HD + i] = num / den;
				}
			}
		}
		mk_bar_sync_all(); // scratch reuse by the next item
	}
}

template <int G>
__device__ void mk_attention_hd(const AttnArgs& a, int HD, float* scratch) {
	switch (HD) {
		case 64: mk_attention<64, G>(a, scratch); break;
		case 128: mk_attention<128, G>(a, scratch); break;
	}
}

__host__ __device__ inline size_t mk_attn_scratch_floats(int HD, int G) {
	const int LPR = HD / 8, RPW = 32 / LPR, NGRP = MK_CW * RPW;
	return (size_t) 3 * NGRP * G + (size_t) MK_CW * G * HD;
}

struct MkArgs {
	const MkPhase* phases;
	int n_phases;
	int NS;               // ring slots PER consumer group
	int slot_bytes;       // bytes per slot (>= 8 rows * U * UB)
	int xb_floats;        // activation staging area (max n over phases, and >= attention scratch)
	unsigned int* gbar;   // grid barrier counter, zero at launch
};

constexpr int MK_GROUPS = 2;              // independent 8-warp consumer groups per CTA (each with its own ring + producer warp)
constexpr int MK_GW = MK_CW / MK_GROUPS;  // warps per group
constexpr int MK_THREADS2 = (MK_CW + MK_GROUPS) * 32;

template <int TYPE>
struct MkCfg2 {
	static constexpr int PPU = UFmt<TYPE>::PPU;
	static constexpr int KW = PPU == 8 ? 4 : 8;   // K-slice warps per group
	static constexpr int RW = MK_GW / KW;         // row groups
	static constexpr int RC = 8;                  // rows per full tile
	static constexpr int R = RC / RW;             // rows per warp
	static constexpr int U = (KW * 32) / PPU > 0 ? (KW * 32) / PPU : 1; // units per stage (one piece per lane per row)
};

__device__ __forceinline__ void mk_group_bar(int group) { asm volatile("bar.sync %0, 256;" ::"r"(2 + group) : "memory"); }

template <int TYPE>
__global__ void __launch_bounds__(MK_THREADS2, 1) layer_megakernel(const MkArgs mk) {
	using F = Fmt<TYPE>;
	using UF = UFmt<TYPE>;
	using C = MkCfg2<TYPE>;
	constexpr int E = F::E, PPU = C::PPU, KW = C::KW, RC = C::RC, R = C::R, U = C::U;
	const int UB = unit_bytes(TYPE);
	const int NS = mk.NS;

	extern __shared__ __align__(128) uint8_t smem[];
	float* xb = reinterpret_cast<float*>(smem);
	uint8_t* ring0 = smem + (((size_t) mk.xb_floats * sizeof(float) + 127) / 128) * 128;
	float* part0 = reinterpret_cast<float*>(ring0 + (size_t) MK_GROUPS * NS * mk.slot_bytes); // [group][2][KW][RC]
	float* s_red = part0 + MK_GROUPS * 2 * KW * RC;                                            // [MK_CW]
	uint64_t* full0 = reinterpret_cast<uint64_t*>(s_red + 32);                                 // [group][NS]
	uint64_t* empty0 = full0 + MK_GROUPS * NS;

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (threadIdx.x == 0) {
		for (int s = 0; s < MK_GROUPS * NS; s++) {
			mbar_init(&full0[s], 1);
			mbar_init(&empty0[s], MK_GW);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	pdl_launch_dependents();
	int tl = -1;
	if (blockIdx.x == 0 && threadIdx.x == 0) tl = tl_begin(400);
	const int n_vcta = (int) gridDim.x * MK_GROUPS; // "virtual CTAs" = consumer groups

	if (warp >= MK_CW) {
		// ===================== producers (one per group): the whole token's weight stream, never blocked by a grid barrier =====
		const int group = warp - MK_CW;
		if (lane == 0) {
			uint8_t* ring = ring0 + (size_t) group * NS * mk.slot_bytes;
			uint64_t* full = full0 + group * NS;
			uint64_t* empty = empty0 + group * NS;
			const int v = (int) blockIdx.x * MK_GROUPS + group;
			int slot = 0, phase = 0;
			for (int ph = 0; ph < mk.n_phases; ph++) {
				const MkPhase& P = mk.phases[ph];
				if (P.kind != MK_MATVEC) continue;
				const int nu = P.a.n / 256, n_tiles = P.n_tiles, kranges = P.kranges, rc = P.pad;
				const int epi = P.a.epi, glu_off = P.a.glu_off;
				const uint8_t* w0 = P.a.w.p0;
				const size_t ws = P.a.w.s0;
				for (int tile = v; tile < n_tiles; tile += n_vcta) {
					const int row0 = tile * rc;
					for (int kr = 0; kr < kranges; kr++) {
						mbar_wait(&empty[slot], phase ^ 1);
						const int u0 = kr * U;
						const int un = min(U, nu - u0);
						const uint32_t bytes = (uint32_t) un * UB;
						mbar_expect_tx(&full[slot], bytes * rc);
						uint8_t* dst = ring + (size_t) slot * mk.slot_bytes;
						for (int r = 0; r < rc; r++) {
							int pr = row0 + r;
							if (epi == EPI_GLU) { // interleave: even r -> W1[o], odd r -> W3[o]
								const int o = (row0 >> 1) + (r >> 1);
								pr = (r & 1) ? glu_off + o : o;
							}
							bulk_g2s(dst + (size_t) r * U * UB, w0 + (size_t) pr * ws + (size_t) u0 * UB, bytes, &full[slot]);
						}
						if (++slot == NS) { slot = 0; phase ^= 1; }
					}
				}
			}
		}
		return;
	}

	// ===================== consumers =====================
	pdl_wait();
	tl_mark(tl, 2);
	const int group = warp / MK_GW, gwarp = warp % MK_GW;
	const int kw = gwarp % KW, rw = gwarp / KW;
	const int my_piece = kw * 32 + lane; // piece index inside a stage row
	const int pu = my_piece / PPU, pp = my_piece % PPU;
	uint8_t* ring = ring0 + (size_t) group * NS * mk.slot_bytes;
	uint64_t* full = full0 + group * NS;
	uint64_t* empty = empty0 + group * NS;
	float* part = part0 + group * 2 * KW * RC;
	const int v = (int) blockIdx.x * MK_GROUPS + group;
	int slot = 0, phase = 0;

	for (int ph = 0; ph < mk.n_phases; ph++) {
		const MkPhase& P = mk.phases[ph];
		int tlp = -1;
		if (blockIdx.x == 0 && threadIdx.x == 0) tlp = tl_begin(410 + (P.kind == MK_ATTN ? 9 : P.a.epi)); // entry = arrival at the barrier
		if (threadIdx.x == 0 && d_timeline.cap >= 100000) { // debug: arrival time of EVERY CTA at every barrier
			const int sl = tl_begin(1000 + ph);
			if (sl >= 0) d_timeline.buf[4 * sl + 2] = blockIdx.x;
		}
		if (ph > 0) mk_grid_barrier(mk.gbar, (unsigned int) ph * gridDim.x);
		tl_mark(tlp, 2);
		if (P.kind == MK_ATTN) {
#ifndef XALM_MK_NO_ATTN
			switch (P.G) {
				case 1: mk_attention_hd<1>(P.at, P.HD, xb); break;
				case 2: mk_attention_hd<2>(P.at, P.HD, xb); break;
				case 4: mk_attention_hd<4>(P.at, P.HD, xb); break;
				case 8: mk_attention_hd<8>(P.at, P.HD, xb); break;
			}
#endif
			tl_mark(tlp, 3);
			continue;
		}
		const MatvecArgs& a = P.a;
		const int n = a.n, nu = n / 256, epi = a.epi, rc = P.pad;
		float* const out = a.out;
		if (epi == EPI_QKV && blockIdx.x == 0 && a.step->kv_sink > 0) {
			const int pairs = a.kv_dim / 2;
			for (int i = threadIdx.x; i < a.step->kv_sink * pairs; i += MK_CW * 32) {
				const int r = i / pairs, p = i % pairs;
				__half2* kp = reinterpret_cast<__half2*>(a.k_cache + (size_t) r * a.kv_dim) + p;
				float2 vv = __half22float2(__ldcg(kp));
				rope_pair(vv.x, vv.y, (2 * p) % a.head_dim, 1, a.rope_freq);
				*kp = __floats2half2_rn(vv.x, vv.y);
			}
		}
		// ---- stage activations (L2 loads: other SMs wrote them in the previous phase), permuted per unit ----
		{
			auto xpos = [](int e) {
				const int u = e >> 8, w = e & 255;
				const int p2 = w / E, i4 = (w % E) >> 2;
				return (u << 8) + ((i4 * PPU + p2) << 2);
			};
			const bool norm = a.norm_w != nullptr;
			float ss = 0.f;
			for (int i = threadIdx.x * 4; i < n; i += MK_CW * 32 * 4) {
				const float4 xvv = ld_cg4(a.x + i);
				*reinterpret_cast<float4*>(xb + xpos(i)) = xvv;
				ss += xvv.x * xvv.x + xvv.y * xvv.y + xvv.z * xvv.z + xvv.w * xvv.w;
			}
			if (norm) {
				ss = warp_sum(ss);
				if (lane == 0) s_red[warp] = ss;
				mk_bar_sync_all();
				float tot = 0.f;
#pragma unroll
				for (int i = 0; i < MK_CW; i++) tot += s_red[i];
				const float scale = 1.0f / sqrtf(tot / (float) n + a.norm_eps);
				for (int i = threadIdx.x * 4; i < n; i += MK_CW * 32 * 4) {
					float4 xvv = *reinterpret_cast<float4*>(xb + xpos(i));
					float4 g;
					if (a.norm_type == XALM_F32) g = ld_act4(reinterpret_cast<const float*>(a.norm_w) + i);
					else {
						const uint2 gv = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(a.norm_w) + i);
						g = make_float4(__uint_as_float(gv.x << 16), __uint_as_float(gv.x & 0xFFFF0000u), __uint_as_float(gv.y << 16),
						                __uint_as_float(gv.y & 0xFFFF0000u));
					}
					xvv.x = xvv.x * scale * g.x; xvv.y = xvv.y * scale * g.y; xvv.z = xvv.z * scale * g.z; xvv.w = xvv.w * scale * g.w; // infer.cpp:233-235
					*reinterpret_cast<float4*>(xb + xpos(i)) = xvv;
				}
			}
			mk_bar_sync_all();
		}
		// activation slice of this lane: fixed for the whole phase when a row is a single K-range
		float xv[E];
		auto load_xv = [&](int u0) {
			const float* xs = xb + (size_t) (u0 + pu) * 256 + pp * 4;
#pragma unroll
			for (int i = 0; i < E; i += 4) {
				const float4 t4 = *reinterpret_cast<const float4*>(xs + i * PPU);
				xv[i] = t4.x; xv[i + 1] = t4.y; xv[i + 2] = t4.z; xv[i + 3] = t4.w;
			}
		};
		const int kranges = P.kranges, n_tiles = P.n_tiles;
		const bool single_range = kranges == 1;
		if (single_range && pu < nu) load_xv(0);
		const int rrows = rc / C::RW; // rows this warp handles per tile (R for full tiles, R/2 for 4-row tiles)

		int tcount = 0;
		for (int tile = v; tile < n_tiles; tile += n_vcta, tcount++) {
			const int row0 = tile * rc;
			const bool reducer = kw == (tcount % KW) && rw == 0;
			// residual: fetch the old activation early so the epilogue does not sit on an L2 round trip
			float xold = 0.f;
			if (epi == EPI_RESIDUAL && reducer && lane < rc) xold = __ldcg(out + row0 + lane);
			f32x2 acc[R];
#pragma unroll
			for (int r = 0; r < R; r++) acc[r] = pack2(0.f, 0.f);
			for (int kr = 0; kr < kranges; kr++) {
				const int u0 = kr * U;
				const bool live = u0 + pu < nu;
				mbar_wait(&full[slot], phase);
				if (live) {
					if (!single_range) load_xv(u0);
					const uint8_t* rows = ring + (size_t) slot * mk.slot_bytes + (size_t) (rw * rrows) * U * UB + (size_t) pu * UB;
#pragma unroll
					for (int r = 0; r < R; r++) {
						if (r < rrows) {
							const typename F::Frag f = UF::load(rows + (size_t) r * U * UB, pp);
							F::fma_chunk(f, xv, acc[r]);
						}
					}
				}
				__syncwarp();
				if (lane == 0) mbar_arrive(&empty[slot]);
				if (++slot == NS) { slot = 0; phase ^= 1; }
			}
			// ---- lanes -> one sum per row; K-slice warps -> shared memory (fixed order) ----
			float* pt = part + (tcount & 1) * (KW * RC);
#pragma unroll
			for (int r = 0; r < R; r++) {
				float lo, hi;
				unpack2(acc[r], lo, hi);
				const float y = warp_sum(lo + hi);
				if (lane == 0 && r < rrows) pt[kw * RC + rw * rrows + r] = y;
			}
			mk_group_bar(group);
			if (reducer) { // rotating reducer warp: lane i owns row i of the tile
				float yv = 0.f;
				if (lane < rc) {
#pragma unroll
					for (int k = 0; k < KW; k++) yv += pt[k * RC + lane];
				}
				const float ynext = __shfl_down_sync(0xffffffffu, yv, 1);
				if (lane < rc) {
					if (epi == EPI_RESIDUAL) {
						out[row0 + lane] = xold + yv;
					} else if (epi == EPI_GLU) {
						if ((lane & 1) == 0) { // (W1[o], W3[o]) sit in adjacent rows of a GLU tile
							const float g = a.act == XALM_SILU ? act_silu(yv) : act_gelu(yv);
							out[(row0 + lane) >> 1] = g * ynext;
						}
					} else if ((lane & 1) == 0) {
						const float y2[2] = {yv, ynext};
						epilogue<2>(a, row0 + lane, y2);
					}
				}
			}
		}
		tl_mark(tlp, 3);
	}
	tl_mark(tl, 3);
}

} // namespace xalm
