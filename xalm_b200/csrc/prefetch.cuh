// prefetch.cuh — keep HBM busy while the compute kernels sit in latency-bound phases.
//
// The in-kernel timeline (profiles/r1_timeline_multikernel.md) shows ~25 us per layer in which no weight byte moves: kernel
// boundaries, activation staging, the attention latency chain.  One warp on a side branch of the token graph walks the
// token's weight matrices in execution order and pulls them into the 126 MB L2 with cp.async.bulk.prefetch.L2, never more
// than `window` bytes ahead of the matvec kernel that is currently running (block 0 of every matvec publishes its index).
// The matvec kernels never wait for it; if it falls behind it skips ahead; a clock-based timeout bounds its life.
#pragma once
#include "common.cuh"

namespace xalm {

struct PrefetchItem {
	const uint8_t* ptr;
	unsigned long long bytes;
	unsigned long long cum_start; // bytes of all earlier items
};

__global__ void __launch_bounds__(32, 1) l2_prefetch_kernel(const PrefetchItem* items, int n_items, const unsigned int* progress,
                                                             unsigned long long window, unsigned long long timeout_ns) {
	const int lane = threadIdx.x;
	const unsigned long long t_start = gtime();
	unsigned long long t_change = t_start;
	unsigned int last_cur = 0xffffffffu;
	constexpr unsigned int CHUNK = 4096; // bytes per lane per request
	for (int k = 0; k < n_items; k++) {
		const PrefetchItem it = items[k];
		for (unsigned long long off = 0; off < it.bytes; off += 32ull * CHUNK) {
			// throttle / skip
			for (;;) {
				const unsigned int cur = *reinterpret_cast<const volatile unsigned int*>(progress);
				if ((int) cur > k) { off = it.bytes; break; } // the consumer is already past this matrix
				const unsigned long long consumed = items[cur < (unsigned) n_items ? cur : n_items - 1].cum_start;
				if (it.cum_start + off < consumed + window) break;
				if (gtime() - t_start > timeout_ns) return;
				if (cur != last_cur) { last_cur = cur; t_change = gtime(); }
				else if (gtime() - t_change > 300000ull) return; // no matvec has started for 300 us: nobody is publishing
				__nanosleep(200);
			}
			if (off >= it.bytes) break;
			const unsigned long long o = off + (unsigned long long) lane * CHUNK;
			if (o < it.bytes) {
				unsigned long long n = it.bytes - o;
				if (n > CHUNK) n = CHUNK;
				n &= ~15ull;
				if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(it.ptr + o), "r"((unsigned int) n) : "memory");
			}
		}
	}
}

} // namespace xalm
