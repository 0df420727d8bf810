// decode_mega.cu — one persistent kernel per decode token (Model::forward, infer.cpp:604-638, batch 1).
//
// Why.  With one kernel per fused op (5 per layer, 162 per Mistral-7B token) the round-1 timeline
// (profiles/r1_timeline_multikernel.md) shows ~36 us of HBM streaming per layer at roofline against 68 us measured: every
// kernel boundary drains the weight stream (ramp-up, tail, dependency flush, activation staging), and the 8-bit consumer
// loop itself was issue/latency-bound.  Here a token is a list of PHASES run by the same resident CTAs, one per SM:
//
//   [norm+QKV+rope+KV] -> [attention] -> [Wo+residual] -> [norm+W1|W3+GLU] -> [W2+residual] -> ... -> [norm+classifier]
//
//   * one PRODUCER warp per CTA walks the token's whole tile list and keeps a shared-memory ring (up to ~170 KB per SM,
//     ~25 MB chip-wide) full with cp.async.bulk copies signalled through mbarriers.  Weights are immutable, so it never
//     waits for a phase hand-off: while the consumers synchronise, stage activations or run attention, the ring fills with
//     the NEXT phases' weights and HBM keeps streaming.  Small phases (Wo, QKV) are consumed straight from shared memory.
//   * eight CONSUMER warps per CTA run the integer-dot core (idp.cuh): activations are staged once per phase as block
//     floating point (three int8 limbs per element), a 32-weight block costs 24 dp4a, ~1.3 issue slots per weight, so the
//     consumers outrun the stream and catch up after every hand-off.
//   * phases are separated by a grid-wide hand-off among the consumer warps: release-add on a counter in global memory,
//     acquire-poll, then the activations written by other SMs are read with ld.global.cg (L2), never through L1.
//   * tile t of a phase runs on CTA (t + tile_off) % grid with tile_off advanced by each phase's remainder, so the odd tile
//     moves around and every SM streams the same number of bytes per token.
//
// Mapping of a matvec phase.  A tile is 8 rows (a RoPE pair / a W1,W3 pair stay adjacent), a ring stage is 8 rows x 16 units
// (4096 elements).  Warp (kw, rw): K-slice kw of 4 (32 blocks of 32 elements, one block per lane) x row group rw of 2
// (4 rows each).  Per stage a lane loads its activation block once (7 shared loads) and walks its 4 rows (2-3 shared loads,
// 24 dp4a and ~8 finishing instructions each).  Lane sums are reduced with a 6-shuffle transposed butterfly, K-slices are
// combined through shared memory in a fixed order (bit-reproducible run to run), a rotating warp runs the epilogue.
#define XALM_SECONDARY_TU
#include "decode_mega.h"

#include <math_constants.h>

#include "idp.cuh"

namespace xalm {

constexpr unsigned int DM_SPIN_LIMIT = 1u << 21; // acquire-polls (~0.5 us each) before a hand-off wait gives up

__device__ __forceinline__ void dm_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ float4 dm_ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// Every wait in this kernel is bounded: a wait that gives up raises the CTA's abort flag (later waits return at once), the
// device-wide abort word gbar[1] (other CTAs stop waiting for this one) and the pinned host word the next synchronising call
// turns into an error — a protocol bug or a lost CTA must not hang the GPU.
__device__ __forceinline__ bool dm_mbar_wait(uint64_t* bar, uint32_t parity, volatile int* s_abort) {
	unsigned int spins = 0;
	for (;;) {
		uint32_t ok;
		asm volatile(
		    "{\n"
		    ".reg .pred p;\n"
		    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
		    "selp.u32 %0, 1, 0, p;\n"
		    "}\n"
		    : "=r"(ok)
		    : "r"(smem_u32(bar)), "r"(parity)
		    : "memory");
		if (ok) return true;
		if ((++spins & 1023u) == 0) {
			if (*s_abort) return false;
			if (spins > (1u << 20)) { *s_abort = 1; return false; }
		}
	}
}

// grid-wide hand-off among the consumer warps of all CTAs: everything this CTA wrote is released, everything the others
// wrote before their arrival is acquired.  `target` = arrivals expected so far (the counter only grows inside a launch).
__device__ __forceinline__ void dm_handoff(unsigned int* ctr, unsigned int target, unsigned int* err, volatile int* s_abort) {
	dm_bar(); // all consumer warps of this CTA have issued their stores
	if (threadIdx.x == 0) {
		__threadfence();
		asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
		unsigned int v, spins = 0;
		for (;;) {
			asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
			if (v >= target) break;
			if ((++spins & 255u) == 0) {
				unsigned int ab;
				asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(ab) : "l"(ctr + 1) : "memory");
				if (ab || *s_abort || spins > DM_SPIN_LIMIT) { // a CTA is missing: report, do not hang the GPU
					*s_abort = 1;
					asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(ctr + 1), "r"(1u) : "memory");
					if (err) *err = 2u;
					break;
				}
			}
		}
	}
	dm_bar();
}

// ---- activation staging: x (global, written by other SMs) [-> rmsnorm] -> block floating point in shared memory ----------
// norm weight chunk i..i+3 as fp32 (F32, or BF16 = bits << 16, types.h:322-325)
__device__ __forceinline__ float4 dm_norm_w4(const MatvecArgs& a, int i) {
	if (a.norm_type == XALM_F32) return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.norm_w) + i);
	const uint2 packed = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(a.norm_w) + i);
	return make_float4(__uint_as_float(packed.x << 16), __uint_as_float(packed.x & 0xFFFF0000u), __uint_as_float(packed.y << 16),
	                   __uint_as_float(packed.y & 0xFFFF0000u));
}
// rmsnorm-fused staging (infer.cpp:224-236): pass 1 pulls x from L2 (all loads of a thread in flight together), accumulates the
// sum of squares and parks the raw values in `stash` (the tail of the activation area, free while n = dim is being staged);
// pass 2 scales, multiplies by the norm weight and quantises.
__device__ __forceinline__ void dm_stage_norm(const MatvecArgs& a, const XqView& v, float* s_red, float* stash) {
	const int n = a.n;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	constexpr int B = 4;
	float ss = 0.f;
	for (int i0 = tid * 4; i0 < n; i0 += B * DM_CW * 32 * 4) {
		float4 xv[B];
#pragma unroll
		for (int c = 0; c < B; c++) {
			const int i = i0 + c * (DM_CW * 32 * 4);
			if (i < n) xv[c] = dm_ld_cg4(a.x + i);
		}
#pragma unroll
		for (int c = 0; c < B; c++) {
			const int i = i0 + c * (DM_CW * 32 * 4);
			if (i < n) {
				ss += xv[c].x * xv[c].x + xv[c].y * xv[c].y + xv[c].z * xv[c].z + xv[c].w * xv[c].w;
				*reinterpret_cast<float4*>(stash + i) = xv[c];
			}
		}
	}
	ss = warp_sum(ss);
	if (lane == 0) s_red[warp] = ss;
	dm_bar();
	float tot = 0.f;
#pragma unroll
	for (int i = 0; i < DM_CW; i++) tot += s_red[i];
	const float scale = 1.0f / sqrtf(tot / (float) n + a.norm_eps); // infer.cpp:229-232
	for (int i = tid * 4; i < n; i += DM_CW * 32 * 4) { // same elements this thread parked above
		const float4 x = *reinterpret_cast<const float4*>(stash + i);
		const float4 g = dm_norm_w4(a, i);
		float4 o;
		o.x = x.x * scale * g.x; // infer.cpp:233-235
		o.y = x.y * scale * g.y;
		o.z = x.z * scale * g.z;
		o.w = x.w * scale * g.w;
		xq_store4(v, i, o, lane);
	}
}

// ---- activation staging: x (global, written by other SMs) [-> rmsnorm] -> block floating point in shared memory ----------
template <bool NORM>
__device__ __forceinline__ void dm_stage(const MatvecArgs& a, uint8_t* xq_base, int xq_cap, float* s_red) {
	const int n = a.n;
	const int tid = threadIdx.x, lane = tid & 31;
	const XqView v = xq_view(xq_base, n);
	if (NORM) {
		dm_stage_norm(a, v, s_red, reinterpret_cast<float*>(xq_base + xq_cap) - n); // host: xq_cap >= xq_bytes(n) + 4 n
	} else {
		constexpr int B = 4; // loads in flight per thread
		for (int i0 = tid * 4; i0 < n; i0 += B * DM_CW * 32 * 4) {
			float4 xv[B];
#pragma unroll
			for (int c = 0; c < B; c++) {
				const int i = i0 + c * (DM_CW * 32 * 4);
				if (i < n) xv[c] = dm_ld_cg4(a.x + i);
			}
#pragma unroll
			for (int c = 0; c < B; c++) {
				const int i = i0 + c * (DM_CW * 32 * 4);
				if (i < n) xq_store4(v, i, xv[c], lane);
			}
		}
	}
	dm_bar();
}

// ---- attention phase: the 8 consumer warps process (kv head, split) items; same math as attn_decode_kernel (attention.cuh) ----
// KVDIV = 2 serves 2 x G query heads per kv head as two "virtual" kv heads of G heads each (G = 8 would need > 200 registers per
// thread; the second pass re-reads the K/V slice from L2).
template <int HD, int G, int KVDIV>
__device__ __forceinline__ void dm_attention(const AttnArgs& a, float* scratch, int first, int stride) {
	constexpr int NW = DM_CW;
	constexpr int LPR = HD / 8, RPW = 32 / LPR, TB = 4, NGRP = NW * RPW;
	float* s_m = scratch;                       // [NGRP][G]
	float* s_l = s_m + NGRP * G;                // [NGRP][G]
	float* s_scale = s_l + NGRP * G;            // [NGRP][G]
	float* s_acc = s_scale + NGRP * G;          // [NW][G][HD]
	__shared__ int s_last;
	const int kv_len = a.kv_len_fixed >= 0 ? a.kv_len_fixed : a.step->kv_len;
	const int slen = attn_split_len(kv_len, a.n_splits, a.min_split);
	const int n_active = (kv_len + slen - 1) / slen;
	const int n_items = a.n_kv_heads * KVDIV * n_active;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int sub = lane / LPR, li = lane % LPR;
	const int kv_stride = a.n_kv_heads * HD;
	const float inv_sqrt = 1.0f / sqrtf((float) HD);
	for (int item = first; item < n_items; item += stride) {
		const int kvh = item / n_active, split = item % n_active; // kvh: virtual kv head (G query heads each)
		const int kvp = kvh / KVDIV;                              // physical kv head
		const int t0 = split * slen, t1 = min(kv_len, t0 + slen);
		float qf[G][8];
#pragma unroll
		for (int g = 0; g < G; g++) {
			const float* qp = a.q + (size_t) (kvh * G + g) * HD + li * 8;
			const float4 u = dm_ld_cg4(qp), v = dm_ld_cg4(qp + 4);
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
		const __half* kbase = a.k_cache + (size_t) kvp * HD + li * 8;
		const __half* vbase = a.v_cache + (size_t) kvp * HD + li * 8;
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
		dm_bar();
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
		dm_bar();
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
			dm_bar();
			if (threadIdx.x == 0) {
				const unsigned int ticket = atomicAdd(&a.tickets[kvh], 1u);
				s_last = ticket == (unsigned int) (n_active - 1);
				if (s_last) a.tickets[kvh] = 0;
			}
			dm_bar();
			if (s_last) { // the last split of this kv head to finish merges all of them, in split order
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
		dm_bar(); // scratch reuse by the next item
	}
}
template <int G, int KVDIV>
__device__ __forceinline__ void dm_attention_hd(const AttnArgs& a, int HD, float* scratch, int first, int stride) {
	if (HD == 64) dm_attention<64, G, KVDIV>(a, scratch, first, stride);
	else dm_attention<128, G, KVDIV>(a, scratch, first, stride);
}

template <int TYPE>
__global__ void __launch_bounds__(DM_THREADS, 1) decode_token_kernel(const DmArgs mk) {
	using F = IdpFmt<TYPE>;
	constexpr int KW = DM_KW, R = DM_R, RC = DM_RC, U = DM_U, UB = F::UB;
	constexpr int ROW_STAGE = U * UB; // bytes of one row inside a ring slot
	const int NS = mk.NS;
	const int G = (int) gridDim.x;

	extern __shared__ __align__(128) uint8_t smem[];
	uint8_t* xq_base = smem;
	uint8_t* ring = smem + mk.xq_cap;
	float* part = reinterpret_cast<float*>(ring + (size_t) NS * mk.slot_bytes); // [2][KW][RC]
	float* s_red = part + 2 * KW * RC;                                           // [16]
	uint64_t* full = reinterpret_cast<uint64_t*>(s_red + 16);
	uint64_t* empty = full + NS;
	__shared__ MatvecArgs s_a; // this phase's arguments (the epilogue reads a dozen fields per tile)
	__shared__ int s_abort;

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (threadIdx.x == 0) {
		s_abort = 0;
		for (int s = 0; s < NS; s++) {
			mbar_init(&full[s], 1);
			mbar_init(&empty[s], DM_CW);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	if (warp == DM_CW) {
		// ===================== producer: the whole token's weight stream, never blocked by a hand-off =====================
		if (lane == 0) {
			int slot = 0, phase = 0;
			for (int ph = 0; ph < mk.n_phases; ph++) {
				const DmPhase& P = mk.phases[ph];
				if (P.kind != DM_MATVEC) continue;
				const int nu = P.a.n / 256, n_tiles = P.n_tiles, kranges = P.kranges;
				const int epi = P.a.epi, glu_off = P.a.glu_off;
				const uint8_t* w0 = P.a.w.p0;
				const size_t ws = P.a.w.s0;
				const int first = ((int) blockIdx.x + G - P.tile_off % G) % G;
				for (int tile = first; tile < n_tiles; tile += G) {
					const int row0 = tile * RC;
					for (int kr = 0; kr < kranges; kr++) {
						if (!dm_mbar_wait(&empty[slot], phase ^ 1, &s_abort)) return;
						const int u0 = kr * U;
						const int un = min(U, nu - u0);
						const uint32_t bytes = (uint32_t) un * UB;
						mbar_expect_tx(&full[slot], bytes * RC);
						uint8_t* dst = ring + (size_t) slot * mk.slot_bytes;
#pragma unroll
						for (int r = 0; r < RC; r++) {
							int pr = row0 + r;
							if (epi == EPI_GLU) { // interleave: even r -> W1[o], odd r -> W3[o]
								const int o = (row0 >> 1) + (r >> 1);
								pr = (r & 1) ? glu_off + o : o;
							}
							bulk_g2s(dst + (size_t) r * ROW_STAGE, w0 + (size_t) pr * ws + (size_t) u0 * UB, bytes, &full[slot]);
						}
						if (++slot == NS) { slot = 0; phase ^= 1; }
					}
				}
			}
		}
		return;
	}

	// ===================== consumers =====================
	const int kw = warp % KW, rw = warp / KW;
	const int hA = (lane >> 2) & 1;
	int slot = 0, phase = 0;
	unsigned long long* tl = (mk.tl && blockIdx.x == 0 && threadIdx.x == 0) ? mk.tl : nullptr;

	for (int ph = 0; ph < mk.n_phases; ph++) {
		const DmPhase& P = mk.phases[ph];
		if (tl) tl[4 * ph] = gtime();
		if (mk.tl && threadIdx.x == 0) mk.tl[(size_t) 4 * mk.tl_phases + (size_t) ph * G + blockIdx.x] = gtime();
		if (ph > 0) dm_handoff(mk.gbar, (unsigned int) ph * (unsigned int) G, mk.err, &s_abort);
		if (tl) tl[4 * ph + 1] = gtime();
		const int first = ((int) blockIdx.x + G - P.tile_off % G) % G;
		if (P.kind == DM_ATTN) {
			float* scratch = reinterpret_cast<float*>(xq_base);
			switch (P.G) {
				case 1: dm_attention_hd<1, 1>(P.at, P.HD, scratch, first, G); break;
				case 2: dm_attention_hd<2, 1>(P.at, P.HD, scratch, first, G); break;
				case 4: dm_attention_hd<4, 1>(P.at, P.HD, scratch, first, G); break;
				case 8: dm_attention_hd<4, 2>(P.at, P.HD, scratch, first, G); break;
			}
			if (tl) { tl[4 * ph + 2] = tl[4 * ph + 1]; tl[4 * ph + 3] = gtime(); }
			continue;
		}
		// this phase's arguments -> shared memory
		{
			const uint32_t* src = reinterpret_cast<const uint32_t*>(&P.a);
			uint32_t* dst = reinterpret_cast<uint32_t*>(&s_a);
			for (int i = threadIdx.x; i < (int) (sizeof(MatvecArgs) / 4); i += DM_CW * 32) dst[i] = src[i];
		}
		dm_bar();
		const MatvecArgs& a = s_a;
		const int n = a.n, nu = n / 256, epi = a.epi;
		float* const out = a.out;
		if (epi == EPI_QKV && blockIdx.x == 0 && a.step->kv_sink > 0) { // attention sinks move on by one position (infer.cpp:416-431)
			const int pairs = a.kv_dim / 2;
			for (int i = threadIdx.x; i < a.step->kv_sink * pairs; i += DM_CW * 32) {
				const int r = i / pairs, p = i % pairs;
				__half2* kp = reinterpret_cast<__half2*>(a.k_cache + (size_t) r * a.kv_dim) + p;
				float2 vv = __half22float2(__ldcg(kp));
				rope_pair(vv.x, vv.y, (2 * p) % a.head_dim, 1, a.rope_freq);
				*kp = __floats2half2_rn(vv.x, vv.y);
			}
		}
		if (a.norm_w != nullptr) dm_stage<true>(a, xq_base, mk.xq_cap, s_red);
		else dm_stage<false>(a, xq_base, mk.xq_cap, s_red);
		if (tl) tl[4 * ph + 2] = gtime();
		const XqView xv = xq_view(xq_base, n);

		const int kranges = P.kranges, n_tiles = P.n_tiles;
		int tcount = 0;
		for (int tile = first; tile < n_tiles; tile += G, tcount++) {
			const int row0 = tile * RC;
			const bool reducer = warp == (tcount % DM_CW);
			// residual: fetch the old activation early so the epilogue does not sit on an L2 round trip
			float xold = 0.f;
			if (epi == EPI_RESIDUAL && reducer && lane < RC) xold = __ldcg(out + row0 + lane);
			float y[R];
#pragma unroll
			for (int r = 0; r < R; r++) y[r] = 0.f;
			for (int kr = 0; kr < kranges; kr++) {
				const int u0 = kr * U;
				const int nb = 8 * min(U, nu - u0); // blocks per row in this stage
				const int b = kw * 32 + lane;
				const bool got = dm_mbar_wait(&full[slot], phase, &s_abort);
				if (got && b < nb) {
					XqBlock xb;
					xq_load(xv, u0 * 8 + b, hA, xb);
					const uint8_t* unit = ring + (size_t) slot * mk.slot_bytes + (size_t) (rw * R) * ROW_STAGE + (size_t) (b >> 3) * UB;
#pragma unroll
					for (int r = 0; r < R; r++) F::block(unit + (size_t) r * ROW_STAGE, b & 7, hA, xb, y[r]);
				}
				__syncwarp();
				if (lane == 0) mbar_arrive(&empty[slot]);
				if (++slot == NS) { slot = 0; phase ^= 1; }
			}
			// ---- lanes -> one sum per row (transposed butterfly: 6 shuffles for 4 rows), K-slices -> shared memory (fixed order) ----
			{
				const bool b4 = lane & 16, b3 = lane & 8;
				float k0 = b4 ? y[2] : y[0], k1 = b4 ? y[3] : y[1];
				const float s0 = b4 ? y[0] : y[2], s1 = b4 ? y[1] : y[3];
				k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
				k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
				float k = b3 ? k1 : k0;
				const float s = b3 ? k0 : k1;
				k += __shfl_xor_sync(0xffffffffu, s, 8);
				k += __shfl_xor_sync(0xffffffffu, k, 4);
				k += __shfl_xor_sync(0xffffffffu, k, 2);
				k += __shfl_xor_sync(0xffffffffu, k, 1);
				if ((lane & 7) == 0) part[(tcount & 1) * (KW * RC) + kw * RC + rw * R + (b4 ? 2 : 0) + (b3 ? 1 : 0)] = k;
			}
			dm_bar();
			if (reducer) { // rotating reducer warp: lane i owns row i of the tile
				const float* pt = part + (tcount & 1) * (KW * RC);
				float yv = 0.f;
				if (lane < RC) {
#pragma unroll
					for (int k = 0; k < KW; k++) yv += pt[k * RC + lane];
				}
				const float ynext = __shfl_down_sync(0xffffffffu, yv, 1);
				if (lane < RC) {
					if (epi == EPI_RESIDUAL) {
						out[row0 + lane] = xold + yv; // infer.cpp:450-452, :492-494
					} else if (epi == EPI_GLU) {
						if ((lane & 1) == 0) { // (W1[o], W3[o]) sit in adjacent rows of a GLU tile
							const float g = a.act == XALM_SILU ? act_silu(yv) : act_gelu(yv);
							out[(row0 + lane) >> 1] = g * ynext; // infer.cpp:470-488
						}
					} else if ((lane & 1) == 0) {
						const float y2[2] = {yv, ynext};
						epilogue<2>(a, row0 + lane, y2);
					}
				}
			}
		}
		if (tl) tl[4 * ph + 3] = gtime();
	}
	if (threadIdx.x == 0 && s_abort && mk.err) *mk.err = 3u;
}

bool dm_supported_type(int type) { return idp_supported(type); }
size_t dm_xq_bytes(int n) { return xq_bytes(n); }

size_t dm_attn_scratch_bytes(int HD, int G) {
	if (G == 8) G = 4; // served as two passes of four heads
	const int LPR = HD / 8, RPW = 32 / LPR, NGRP = DM_CW * RPW;
	return ((size_t) 3 * NGRP * G + (size_t) DM_CW * G * HD) * sizeof(float);
}

size_t dm_fixed_smem(size_t xq_cap, int NS) {
	return xq_cap + 2 * DM_KW * DM_RC * sizeof(float) + 16 * sizeof(float) + 2 * (size_t) NS * sizeof(uint64_t) + 128;
}

template <int TYPE>
static cudaError_t dm_launch_typed(const DmArgs& args, int grid, size_t smem, cudaStream_t s, bool coop) {
	auto kern = decode_token_kernel<TYPE>;
	cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
	if (e != cudaSuccess) return e;
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(grid);
	cfg.blockDim = dim3(DM_THREADS);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeCooperative; // every CTA must be resident: the hand-offs spin on one another
	attr[0].val.cooperative = coop ? 1 : 0;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, kern, args);
}

cudaError_t dm_launch(int type, const DmArgs& args, int grid, size_t smem, cudaStream_t s, bool coop) {
	switch (type) {
		case XALM_Q8_0: return dm_launch_typed<XALM_Q8_0>(args, grid, smem, s, coop);
		case XALM_Q8: return dm_launch_typed<XALM_Q8>(args, grid, smem, s, coop);
		case XALM_Q4_0: return dm_launch_typed<XALM_Q4_0>(args, grid, smem, s, coop);
		case XALM_Q4_1: return dm_launch_typed<XALM_Q4_1>(args, grid, smem, s, coop);
		case XALM_Q5_0: return dm_launch_typed<XALM_Q5_0>(args, grid, smem, s, coop);
		case XALM_Q5_1: return dm_launch_typed<XALM_Q5_1>(args, grid, smem, s, coop);
	}
	return cudaErrorInvalidValue;
}

} // namespace xalm
