// sweep_common.cuh -- what the block-sweep kernels for long triples share (K2L fast_long.cuh, K2S fast_stream.cuh):
// the task descriptor and kernel arguments of the single-triple sweep, the claim-position -> descriptor mapping, and the
// cp.async / 128-bit store helpers (plain copies in the host emulation).  The block sweep is the GPU form of the RTL's
// slicing method (src/TriAlign_1cyc.v:127-140,289-335): the (j,k) plane is cut into blocks, every block is swept over
// all of A, and the faces between blocks are parked in memory -- face SRAMs there, HBM/L2 here.
#pragma once

#include "fast_core.cuh"

namespace trialign {
namespace fast {


// Everything a CTA needs to know about one task (CTA-uniform, lives in shared memory).
struct SweepTask {
    int valid;
    int bj, bkp;                  // block row, block column (LONG: column PAIR)
    long long pair;               // unused (kept for layout compatibility of the emulation)
    uint32_t *fj_w, *fk_w;        // faces this task writes (nullptr: nobody needs them)
    const uint32_t *fj_r, *fk_r;  // faces this task reads  (nullptr: zero border plane)
    int top_task, left_task;      // LONG: tasks whose progress gates fj_r / fk_r (-1: none)
    int self_task;                // LONG: index into progress[]
    int final_here;               // this task holds cell (lb,lc)
};

struct SweepArgs {
    const uint8_t *a, *b, *c;     // one triple
    long long n;
    int la, lb, lc;
    int nbj, nbkp;                // blocks per problem: rows, columns (LONG: column pairs)
    long long tasks_total;        // LONG: nbj*nbkp
    long long pairs_total;        // BATCH: ceil(n/2)
    uint32_t *face_j, *face_k;    // ring storage
    long long fj_slot_words, fk_slot_words;
    int fj_ring, fk_ring;         // LONG: global ring sizes; BATCH: per-CTA ring sizes
    int gate;                     // LONG: steps between two progress releases in steady state (power of two)
    int rgate;                    // K2L column pipeline: steps between two system-scope releases to the neighbour GPU
    int gate_tid;                 // K2L: the thread that polls / releases the progress counters
    int early;                    // K2L: 1 = fine publication rhythm during a task's first steps (experiments)
    int pingpong;                 // K2L: middle phase unrolled by two over two alternating register sets
    int *progress;                // LONG: [tasks_total] indexed by task id bj + nbj*bkp, zero-initialised
    const int *order;             // LONG: claim position -> task id (sorted by earliest possible start, long2_order_key)
    const int *pos_of;            // LONG: task id -> claim position
    unsigned long long *counter;  // LONG: task counter, zero-initialised
    // i-slab pipeline (LONG, several GPUs): this launch computes planes i_begin+1 .. i_begin+la of the triple (a already
    // points at symbol i_begin).  The plane-i_begin state of every task comes from the previous slab (iface_in, on
    // this GPU) and the plane-(i_begin+la) state goes to the next slab's GPU (iface_out, possibly a peer pointer).
    // Layout: [task][all,wX,u1,u3][cell 0..15][thread 0..255] packed words, re-based so that the reader sees them as
    // its plane 0.  Flags: one int per task, 1 = state complete (release/acquire at system scope).
    const uint32_t *iface_in;
    uint32_t *iface_out;
    const int *iface_in_flag;
    int *iface_out_flag;
    // column pipeline (K2L, several GPUs): this launch sweeps a contiguous range of block-column pairs of the triple (c
    // already points at the range's first column, lc = its columns).  The K faces crossing the range's left edge come
    // from the left neighbour GPU (cface_in, cprog_in: on this GPU, written by the peer), those crossing its right edge go
    // to the right neighbour (cface_out, cprog_out: peer pointers): one face slot and one progress word per block row.
    // Progress words are tagged (epoch << 20 | steps, 0xFFFFF = done), so that runs repeat without clearing.
    const uint32_t *cface_in;
    const int *cprog_in;
    uint32_t *cface_out;
    int *cprog_out;
    int last_slab;                // emit the score (the launch holds cell (la,lb,lc) of the whole triple)
    int epoch;                    // value the hand-over flags take in this run (lets runs repeat without clearing)
    int32_t *out;                 // BATCH: scores[n]; LONG: out[0] = score
    int32_t *out_states;          // LONG: 7 final-cell states (or nullptr)
    int base;                     // potential base
    int check_lo, check_hi;       // range check (CHECK kernels): every sampled `all` must lie in [check_lo, check_hi]
    int *viol_flag;               // set non-zero when the range check fails
};

constexpr int kProgressDone = 1 << 30;
constexpr int kEarlySteps = 64, kEarlyGate = 2;  // K2L option k2_early: fine publication rhythm at a task's start
constexpr int kCertifiedAll = 1 << 30;

// claim position -> descriptor.  Face ring slots are indexed by the claim position.
template <class C>
TA_HD void sweep_describe_long(const SweepArgs &A, long long pos, SweepTask &T)
{
    const int t = A.order[pos];
    T.valid = 1;
    T.pair = 0;
    T.bkp = t / A.nbj;
    T.bj = t % A.nbj;
    T.self_task = t;
    T.top_task = T.bj > 0 ? t - 1 : -1;
    T.left_task = T.bkp > 0 ? t - A.nbj : -1;
    T.fj_w = (T.bj + 1 < A.nbj) ? A.face_j + (pos % A.fj_ring) * A.fj_slot_words : nullptr;
    T.fk_w = (T.bkp + 1 < A.nbkp) ? A.face_k + (pos % A.fk_ring) * A.fk_slot_words : nullptr;
    T.fj_r = T.bj > 0 ? A.face_j + (A.pos_of[t - 1] % A.fj_ring) * A.fj_slot_words : nullptr;
    T.fk_r = T.bkp > 0 ? A.face_k + (A.pos_of[t - A.nbj] % A.fk_ring) * A.fk_slot_words : nullptr;
    T.final_here = (T.bj == (A.lb - 1) / C::BJ) && (T.bkp == (A.lc - 1) / C::KSPAN);
}

// ---- global <-> shared copies of face rows: cp.async on the device, plain copies in the host emulation ----
#if defined(__CUDA_ARCH__)
TA_HD void ta_copy16(uint32_t *dst_smem, const uint32_t *src_gmem)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src_gmem) : "memory");
}
TA_HD void ta_copy4(uint32_t *dst_smem, const uint32_t *src_gmem)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src_gmem) : "memory");
}
TA_HD void ta_copy_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
TA_HD uint32_t ta_prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
#else
TA_HD void ta_copy16(uint32_t *d, const uint32_t *s) { d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3]; }
TA_HD void ta_copy4(uint32_t *d, const uint32_t *s) { d[0] = s[0]; }
TA_HD void ta_copy_wait() {}
TA_HD uint32_t ta_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int n = 0; n < 4; ++n) r |= (uint32_t)((v >> (8 * ((sel >> (4 * n)) & 7))) & 0xFF) << (8 * n);
    return r;
}
#endif

// 16-byte store of four packed words (dst is 16-byte aligned: face rows are 64 / 68 words, tiles start at multiples of 4)
#if defined(__CUDA_ARCH__)
TA_HD void ta_store4(uint32_t *dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    *reinterpret_cast<uint4 *>(dst) = make_uint4(a, b, c, d);
}
#else
TA_HD void ta_store4(uint32_t *dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    dst[0] = a; dst[1] = b; dst[2] = c; dst[3] = d;
}
#endif

// low half <- high half of x ; high half <- low half of y
TA_HD uint32_t hi_lo(uint32_t x, uint32_t y) { return ta_prmt(x, y, 0x5432); }

}  // namespace fast
}  // namespace trialign
