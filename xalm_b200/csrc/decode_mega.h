// decode_mega.h — seam between xalm_cuda.cu (model handle, graphs) and decode_mega.cu (one persistent kernel per decode token).
#pragma once
#include "attention.cuh"
#include "matvec.cuh"

namespace xalm {

enum { DM_MATVEC = 0, DM_ATTN = 1 };

// One phase of a token: a fused matvec (norm+QKV+rope+KV, Wo+residual, norm+W1|W3+GLU, W2+residual, norm+classifier) or the
// decode attention of a layer.  Phases run in order on the same resident CTAs.
struct DmPhase {
	int kind;
	int n_tiles;   // matvec: virtual rows / 8
	int kranges;   // matvec: ring stages per tile
	int tile_off;  // matvec/attention: work item t runs on CTA (t + tile_off) % grid (rotated so the odd item moves around)
	MatvecArgs a;  // matvec: weights, epilogue kind, norm weights, KV cache rows, plain outputs (x, logits)
	AttnArgs at;   // attention: cache pointers, split geometry
	int G, HD;     // attention: query heads per kv head, head dim
};

struct DmArgs {
	const DmPhase* phases;
	int n_phases;
	int NS;                     // ring slots
	int slot_bytes;             // 8 rows x 16 units
	int xq_cap;                 // bytes reserved for the staged activations / attention scratch
	unsigned int* gbar;         // [0] = hand-off counter, [1] = device-wide abort word; zero at launch
	unsigned int* err;          // pinned host word: set when a wait gives up
	const StepParams* step;     // this token's scalars
	const float* rope_freq;     // (head_dim / 2,)
	int head_dim;
	int quiet;                  // 1: the producer holds its copies back while the consumers are between phases
	unsigned long long* tl;     // optional timeline: tl_phases x 4 stamps of CTA 0, then tl_phases x grid phase-entry stamps (or nullptr)
	int tl_phases;              // phases the timeline buffer was sized for
};

constexpr int DM_CW = 8;                       // consumer warps
constexpr int DM_THREADS = (DM_CW + 1) * 32;   // + one producer warp
constexpr int DM_KW = 4, DM_RW = 2, DM_R = 4;  // K-slice warps x row groups, rows per warp
constexpr int DM_RC = DM_RW * DM_R;            // rows per tile
constexpr int DM_U = 16;                       // units (of 256 elements) per ring stage
constexpr int DM_MAX_SPLITS = 32;              // attention splits per kv head the distributed merge takes
constexpr int DM_MAX_HD = 128;

bool dm_supported_type(int type);
size_t dm_attn_scratch_bytes(int HD, int G);
// shared-memory bytes of a staged activation vector of n elements (three int8 limb planes + per-block sums and scale)
size_t dm_xq_bytes(int n);
// fixed shared-memory bytes besides the ring (activations, partial sums, barriers)
size_t dm_fixed_smem(size_t xq_cap, int NS);
// coop: cooperative launch (the driver refuses the launch instead of letting it deadlock when the CTAs cannot all be resident)
cudaError_t dm_launch(int type, const DmArgs& args, int grid, size_t smem, cudaStream_t s, bool coop);

} // namespace xalm
