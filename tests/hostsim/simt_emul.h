// TEST TOOL: a 32-lane warp emulated with 32 host threads, for the ASTC_HOSTSIM_LANES32 build of tests/hostsim.
//
// Every lane of the simulated warp is an OS thread running the device source; the CUDA warp primitives the source uses
// (__shfl_*_sync, __ballot/__any/__all/__match_any_sync, __syncwarp, and the CTA barrier of a one-warp CTA) are collectives
// over those threads: write a slot, meet at a barrier, read, meet again. All calls in the device source use the full mask,
// so every collective waits for all 32 lanes - a lane that skips one (a collective under lane-divergent control flow) hangs
// the simulation, which the barrier turns into an abort after a timeout. That is the point of this build: the one-lane
// simulation cannot see lane-parallel mistakes (wrong shuffle partner, missing __syncwarp, results that are not uniform
// across the warp); this one runs the same instruction-level protocol as the GPU, without a GPU.
#pragma once
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace simt {
constexpr int LANES = 32;

struct Warp {
	std::atomic<unsigned int> arrived{0};
	std::atomic<unsigned int> generation{0};
	uint64_t slot[LANES];
};
static Warp g_warp;
static thread_local int t_lane = 0;

// sense-reversing barrier over the 32 lane threads (more threads than cores is the normal case: yield while waiting)
static inline void barrier() {
	unsigned int gen = g_warp.generation.load(std::memory_order_acquire);
	if (g_warp.arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == (unsigned int)LANES) {
		g_warp.arrived.store(0, std::memory_order_relaxed);
		g_warp.generation.store(gen + 1, std::memory_order_release);
		return;
	}
	unsigned int spins = 0;
	auto t0 = std::chrono::steady_clock::now();
	while (g_warp.generation.load(std::memory_order_acquire) == gen) {
		std::this_thread::yield();
		if ((++spins & 0xFFFFF) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(60)) {
			fprintf(stderr, "simt_emul: lane %d waited 60 s at a warp collective - some lane never arrived (collective under divergent control flow?)\n", t_lane);
			abort();
		}
	}
}

template <typename T> static inline T exchange(T v, int src_lane) {
	static_assert(sizeof(T) <= 8, "shuffle payload");
	uint64_t bits = 0;
	memcpy(&bits, &v, sizeof(T));
	g_warp.slot[t_lane] = bits;
	barrier();
	uint64_t r = g_warp.slot[src_lane & (LANES - 1)];
	barrier();
	T out;
	memcpy(&out, &r, sizeof(T));
	return out;
}

static inline unsigned int ballot(bool p) {
	g_warp.slot[t_lane] = p ? 1u : 0u;
	barrier();
	unsigned int m = 0;
	for (int l = 0; l < LANES; l++) {
		m |= (unsigned int)(g_warp.slot[l] & 1u) << l;
	}
	barrier();
	return m;
}

// run fn(lane) on 32 threads, one per lane
template <typename F> static inline void run_warp(F fn) {
	g_warp.arrived.store(0);
	std::thread th[LANES];
	for (int l = 0; l < LANES; l++) {
		th[l] = std::thread([l, &fn]() {
			t_lane = l;
			fn(l);
		});
	}
	for (int l = 0; l < LANES; l++) {
		th[l].join();
	}
}
}      // namespace simt

// ---- the CUDA names the device source uses --------------------------------------------------------------------
static inline void __syncwarp() { simt::barrier(); }
static inline void __syncthreads() { simt::barrier(); }                  // the simulated CTA is one warp
template <typename T> static inline T __shfl_sync(unsigned int, T v, int src_lane) { return simt::exchange(v, src_lane); }
template <typename T> static inline T __shfl_xor_sync(unsigned int, T v, int lane_mask) { return simt::exchange(v, simt::t_lane ^ lane_mask); }
template <typename T> static inline T __shfl_up_sync(unsigned int, T v, unsigned int delta) {
	int src = simt::t_lane - (int)delta;
	return simt::exchange(v, src < 0 ? simt::t_lane : src);
}
static inline unsigned int __ballot_sync(unsigned int, int p) { return simt::ballot(p != 0); }
static inline int __any_sync(unsigned int, int p) { return simt::ballot(p != 0) != 0; }
static inline int __all_sync(unsigned int, int p) { return simt::ballot(p != 0) == 0xFFFFFFFFu; }
static inline int __syncthreads_or(int p) { return simt::ballot(p != 0) != 0; }
template <typename T> static inline unsigned int __match_any_sync(unsigned int, T key) {
	uint64_t bits = 0;
	memcpy(&bits, &key, sizeof(T));
	simt::g_warp.slot[simt::t_lane] = bits;
	simt::barrier();
	unsigned int m = 0;
	for (int l = 0; l < simt::LANES; l++) {
		m |= (unsigned int)(simt::g_warp.slot[l] == bits) << l;
	}
	simt::barrier();
	return m;
}
static inline int __popc(unsigned int v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
// (one simulated warp: the queue counters are only ever touched by lane 0 of that warp)
static inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
