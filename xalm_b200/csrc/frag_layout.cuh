// frag_layout.cuh — the "fragment tile" device layout of the integer block formats (Q8_0, Q4_0, Q4_1, Q5_0, Q5_1 of
// quants.py:281-464 and Q8 = int8 x 1/100 of types.h:423-424) that the tensor-core decode matvec (matvec_mma.cuh) streams.
//
// The matrix is cut into tiles of 16 rows; a tile is a run of n/32 RECORDS, one per 32-element quant block column, and a
// record holds the 16 rows' quants of that block in the register order of the A operand of mma.sync.m16n8k32 (lane t = 4g + tig
// owns rows g and g + 8, elements 4 tig .. 4 tig + 3 and 16 + 4 tig .. of the block), followed by the 16 scales (and minimums /
// fifth bits).  Same byte count as on disk:
//
//   Q8_0  [32 lanes x 16 B: a0 a1 a2 a3, quants stored + 128][8 x {d[g], d[g+8]} f16]                          544 B
//   Q8    [32 lanes x 16 B]                                                                                   512 B
//   Q4_0  [32 lanes x 8 B: w0 w1; low nibble = row g, high nibble = row g + 8][8 x {d[g], d[g+8]}]            288 B
//   Q4_1  [32 x 8 B][8 x {d[g], d[g+8], m[g], m[g+8]}]                                                        320 B
//   Q5_0  [32 x 8 B][32 x u16 fifth bits: bit 4j + e -> byte e of a_j][8 x {d[g], d[g+8]}]                    352 B
//   Q5_1  [32 x 8 B][32 x u16][8 x {d, d, m, m}]                                                              384 B
//
// so ONE bulk copy fetches any run of records of a tile, a lane gets its whole A fragment with one 128-bit (or 64-bit) shared
// load, conflict-free, and the two rows of a GLU pair (W1 row o, W3 row o) sit in the same lane: a gate|up matrix is stored
// with virtual rows [16T, 16T+8) = W1 rows [8T, 8T+8) and [16T+8, 16T+16) = W3 rows [8T, 8T+8).
#pragma once
#include "formats.cuh"

namespace xalm {

__host__ __device__ inline int frag_record_bytes(int t) {
	switch (t) {
		case XALM_Q8_0: return 544;
		case XALM_Q8: return 512;
		case XALM_Q4_0: return 288;
		case XALM_Q4_1: return 320;
		case XALM_Q5_0: return 352;
		case XALM_Q5_1: return 384;
	}
	return 0; // not a fragment-layout format
}
__host__ __device__ inline int frag_quant_bytes(int t) { return (t == XALM_Q8_0 || t == XALM_Q8) ? 512 : 256; }

// physical row of the uploaded (stacked) matrix -> virtual row of the tile order
__host__ __device__ inline int frag_virtual_row(int glu_half, int r) {
	if (!glu_half) return r;
	const int o = r < glu_half ? r : r - glu_half;
	return (o / 8) * 16 + (r < glu_half ? 0 : 8) + o % 8;
}

// ---- element access (the batched prefill's dequantiser, tests): 8 consecutive elements k0 .. k0+7 (k0 % 8 == 0) of physical row r
__device__ __forceinline__ void frag_decode8(const WMat& w, int t, int r, int k0, float (&v)[8]) {
	const int vr = frag_virtual_row(w.glu_half, r);
	const int T = vr >> 4, g = vr & 7, hi = (vr >> 3) & 1;
	const int nb = w.n / 32, blk = k0 / 32, kk0 = k0 % 32;
	const int rb = frag_record_bytes(t), qb = frag_quant_bytes(t);
	const uint8_t* rec = w.p0 + ((size_t) T * nb + blk) * rb;
	const int half = kk0 / 16;
	const int j = half * 2 + hi;
	float d = 0.01f, mn = 0.f; // Q8: fixed scale 1/100
	if (t == XALM_Q8_0 || t == XALM_Q4_0) d = f16_bits_to_f32(*reinterpret_cast<const uint16_t*>(rec + qb + g * 4 + hi * 2));
	if (t == XALM_Q5_0) d = f16_bits_to_f32(*reinterpret_cast<const uint16_t*>(rec + qb + 64 + g * 4 + hi * 2));
	if (t == XALM_Q4_1 || t == XALM_Q5_1) {
		const uint8_t* sp = rec + qb + (t == XALM_Q5_1 ? 64 : 0) + g * 8 + hi * 2;
		d = f16_bits_to_f32(*reinterpret_cast<const uint16_t*>(sp));
		mn = f16_bits_to_f32(*reinterpret_cast<const uint16_t*>(sp + 4));
	}
#pragma unroll
	for (int i = 0; i < 8; i++) {
		const int kk = kk0 + i, tig = (kk % 16) / 4, e = kk % 4, lane = g * 4 + tig;
		if (t == XALM_Q8_0 || t == XALM_Q8) {
			const int q = (int) rec[lane * 16 + j * 4 + e] - 128;
			v[i] = __fmul_rn(d, (float) q);
			continue;
		}
		const uint8_t byte = rec[lane * 8 + half * 4 + e];
		int q = hi ? (byte >> 4) : (byte & 15);
		if (t == XALM_Q5_0 || t == XALM_Q5_1) {
			const uint32_t hb = *reinterpret_cast<const uint16_t*>(rec + 256 + lane * 2);
			q |= (int) ((hb >> (4 * j + e)) & 1u) << 4;
		}
		switch (t) {
			case XALM_Q4_0: v[i] = __fmul_rn(d, (float) (q - 8)); break;
			case XALM_Q5_0: v[i] = __fmul_rn(d, (float) (q - 16)); break;
			default: v[i] = __fadd_rn(__fmul_rn(d, (float) q), mn); break; // Q4_1, Q5_1
		}
	}
}

#ifndef XALM_SECONDARY_TU // non-template kernel: defined once, in xalm_cuda.cu's translation unit
// on-disk rows of one uploaded piece (rows [dst_row, dst_row + rows) of the stacked matrix) -> fragment tiles.
// One thread per (tile, g, block): it owns both rows g and g + 8 of its tile, whose nibbles / fifth bits share bytes — the row that is
// not part of this piece (the other half of a gate|up pair arrives with its own upload) is left as it is (read-modify-write).
__global__ void repack_frag_kernel(int t, const uint8_t* __restrict__ raw, size_t raw_stride, int dst_row, int rows, int total_rows, int n,
                                   int glu_half, uint8_t* __restrict__ dst) {
	TypeInfo ti;
	type_info(t, &ti);
	const int nb = n / 32, rb = frag_record_bytes(t), qb = frag_quant_bytes(t);
	const size_t total = (size_t) (total_rows / 16) * 8 * nb;
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x) {
		const int b = (int) (i % nb);
		const int g = (int) ((i / nb) % 8);
		const int T = (int) (i / nb / 8);
		uint8_t* rec = dst + ((size_t) T * nb + b) * rb;
		for (int hi = 0; hi < 2; hi++) {
			const int R = glu_half ? (hi ? glu_half + T * 8 + g : T * 8 + g) : T * 16 + hi * 8 + g; // physical row of the stacked matrix
			const int li = R - dst_row;
			if (li < 0 || li >= rows) continue;
			const uint8_t* s = raw + (size_t) li * raw_stride + (size_t) b * (ti.block == 1 ? 32 * ti.bytes : ti.bytes);
			// disk block -> (scale, min, fifth bits, quants)
			int so = 0;
			const uint8_t *dp = nullptr, *mp = nullptr, *hp = nullptr;
			if (t != XALM_Q8) { dp = s; so = 2; }
			if (t == XALM_Q4_1 || t == XALM_Q5_1) { mp = s + 2; so = 4; }
			if (t == XALM_Q5_0 || t == XALM_Q5_1) { hp = s + so; so += 4; }
			const uint8_t* q = s + so;
			if (t == XALM_Q8_0 || t == XALM_Q8) {
				for (int kk = 0; kk < 32; kk++) {
					const int tig = (kk % 16) / 4, e = kk % 4, j = (kk / 16) * 2 + hi;
					rec[(g * 4 + tig) * 16 + j * 4 + e] = q[kk] ^ 0x80;
				}
			} else {
				const uint32_t qh = hp ? (uint32_t) hp[0] | ((uint32_t) hp[1] << 8) | ((uint32_t) hp[2] << 16) | ((uint32_t) hp[3] << 24) : 0u;
				for (int kk = 0; kk < 32; kk++) {
					const int tig = (kk % 16) / 4, e = kk % 4, half = kk / 16, j = half * 2 + hi;
					const uint8_t nib = half ? (q[kk - 16] >> 4) : (q[kk] & 15);
					uint8_t* p = rec + (g * 4 + tig) * 8 + half * 4 + e;
					*p = hi ? (uint8_t) ((*p & 0x0F) | (nib << 4)) : (uint8_t) ((*p & 0xF0) | nib);
					if (hp) {
						uint16_t* hw = reinterpret_cast<uint16_t*>(rec + 256 + (g * 4 + tig) * 2);
						const uint16_t bit = (uint16_t) (1u << (4 * j + e));
						*hw = ((qh >> kk) & 1u) ? (uint16_t) (*hw | bit) : (uint16_t) (*hw & ~bit);
					}
				}
			}
			if (t == XALM_Q8) continue;
			uint8_t* sp = rec + qb + (hp ? 64 : 0);
			if (mp) {
				sp += g * 8 + hi * 2;
				sp[0] = dp[0]; sp[1] = dp[1]; sp[4] = mp[0]; sp[5] = mp[1];
			} else {
				sp += g * 4 + hi * 2;
				sp[0] = dp[0]; sp[1] = dp[1];
			}
		}
	}
}
#endif // XALM_SECONDARY_TU

} // namespace xalm
