// Fused persistent engine for the structured path (n >= 12), sm_100a.
//
// One product-formula step (diffqc.cc:155-164) = diagonal phase D(k) then X rotations on every
// qubit.  Qubits are split into two sets by physical bit position, L = bits [0,10) and
// H = bits [10,n).  Because the X rotations of one step commute with each other and D is
// elementwise, the work is regrouped into PASSES that each touch one set only:
//
//     pass p :  [ mixer S_p of step p ]  ->  D(p+1)  ->  [ mixer S_p of step p+1 ]        S_p alternates L,H
//
// so a trajectory of N steps costs N+1 passes = ONE global read+write of the state per step.
// Inside a pass a CTA owns a tile of 2^12 amplitudes (64 KiB of shared memory); every thread
// keeps 32 amplitudes in registers and applies 5 qubits' butterflies per register round:
//     TMA tile load -> outer-A (5 "K" bits)  -> smem -> inner (5 "J" bits, phase, 5 "J" bits)
//     -> smem -> outer-B (5 "K" bits) -> global / fused energy reduction.
// Tiles arrive AND leave by TMA (cp.async.bulk.tensor, one instruction per 64 KiB tile): they land in the
// hardware 128-byte swizzle and complete on an mbarrier; results are written back into the same buffer and
// stored with one bulk-tensor store.  The K / J bit sets of both pass types are chosen so that this one swizzle
// is bank-conflict-free in both register rounds (see Geo<>).  A CTA (one per SM) runs two independent TEAMS of
// four warps that share THREE tile buffers: while a team works in one buffer, the third one receives its next
// tile or drains the other team's store, so neither the load latency nor the store ever stalls a warp.
// Rotations use the scaled form a' = a - i tan(theta) b (2 FMA per amplitude per qubit); the product
// of cosines is folded into the phase tables.  The diagonal phase is never computed per amplitude
// with sincos: per thread it is  TC[column] * TKK[K bits] * prod XK_m  (base phase) and then a
// product-state doubling over the 5 register bits with per-bit factors F_k (31 + 32 complex
// multiplies per 32 amplitudes).  Tables are built per (trajectory, pass) by a setup kernel.
//
// The kernel is persistent: CTAs pull (pass, ket, tile) items from an atomic counter in pass-major
// order; an item waits until all tiles of the previous pass of ITS ket are done (per-ket counter,
// release/acquire through L2).  Several kets per launch keep every SM busy across pass boundaries,
// and the ket group is sized to stay resident in the 126 MB L2 between passes.
#include <algorithm>
#include <cuda.h>
#include <math.h>
#include <string.h>
#include "ising.cuh"

namespace dq {
namespace fused {

constexpr int kTileBits = 12;
constexpr int kTile = 1 << kTileBits;      // amplitudes per tile
constexpr int kTeamThreads = 128;          // kTile / 32: one tile is processed by a team of four warps
constexpr int kTeams = 2;                  // teams per CTA
constexpr int kThreads = kTeams * kTeamThreads;
constexpr int kBufs = 3;                   // tile buffers per CTA (one per team + one in flight)
constexpr int kRegs = 32;                  // amplitudes per thread
constexpr int kMaxNbr = 3;                 // neighbour bits per register qubit held in tables
constexpr int kMaxPairs = 128;
constexpr int kMaxGroup = 128;             // kets per launch (their descriptors are staged in shared memory)
#ifndef DQ_EXP
#define DQ_EXP 0      // timing experiments only: 1 no FP64 math, 2 no smem exchange, 4 no global traffic
#endif
#ifndef DQ_TRACE
#define DQ_TRACE 0    // 1: lane 0 of every warp records clock64() at phase boundaries of each item (debug)
#endif
#ifndef DQ_STORE_EVICT_LAST
#define DQ_STORE_EVICT_LAST 1
#endif
#ifndef DQ_PACE_TIGHT
#define DQ_PACE_TIGHT 0   // experiment: the leader waits for the follower's mid-tile barrier instead of its J1 mark (measured: -2 %)
#endif
#ifndef DQ_PACE_EARLY
#define DQ_PACE_EARLY 0   // experiment: the follower team starts after the leader's outer-A round instead of after its J1 rotations
#endif
#ifndef DQ_CTAS_PER_SM
#define DQ_CTAS_PER_SM 1
#endif
#ifndef DQ_WS_PIPE
#define DQ_WS_PIPE 0           // experiment (negative: 79.3 vs 84.4 samples/s at n = 20): exchange stores interleaved with the last butterfly stage of the producing run -- issue is in order, so a store that finds the LSU queue full holds back the DFMAs behind it, and the hot code grows by 560 instructions
#endif
#ifndef DQ_WS_DIRECT_STORE
#define DQ_WS_DIRECT_STORE 0   // experiment (negative: 81.1 vs 84.4 samples/s at n = 20; with a fence per consumer thread 76.7): finished tiles go from the registers straight to global memory instead of shared memory + bulk store -- the 16-byte global stores take the same LSU wavefronts as the shared-memory stores they replace, and the asynchronous bulk store is lost
#endif


enum : int { F_ENERGY = 2, F_STORE = 4, F_SCALED = 8 };   // F_SCALED: setup jobs only (lifting-form tables for this trajectory)

struct __align__(16) PassStep {
    double2 tkk[32];                // base phase over the 5 K bits (includes constant + cos scale)
    double2 aj[32];                 // J-internal pairs, relative to j = 0
    double2 xk[5][8];               // K bit m x pattern of its column neighbours
    double2 fj[5][8];               // J bit k x pattern of its non-J neighbours: exp(+2i w_k)
    double2 rot[4][5];              // (cos, sin) or (1, tan) per slot {KA, J1, J2, KB} and bit
    int flags;
    int type;
    unsigned long long tc_offset;   // first entry of this pass-step's column table
};

// Launch-constant geometry of one pass type; lives in kernel parameters (constant bank).
struct TypeGeom {
    int a, lowmask;                 // tile bit t -> physical: t < a ? t : start + (t - a)   (H: start = 10; L: a = start = 12, identity)
    int start;
    int tid_lo_bits, high_end;
    int fj_pos[5][kMaxNbr], fj_msk[5][kMaxNbr];
    int xk_pos[5][kMaxNbr], xk_msk[5][kMaxNbr];
    int offK[kRegs], offJ[kRegs];   // physical offset (amplitudes) of register j in the outer / inner round
};

// Host-side plan of a pass type (superset of TypeGeom; the setup kernel reads it from global memory).
struct TypePlan {
    TypeGeom g;
    int start, spare_shift;
    int jq[5], kq[5];               // x-angle column of the qubit on that register bit, -1 = spectator
    int jq_all[5], kq_all[5];       // qubit whose index bit sits on that register position, rotated by this pass type or not
    int colq[16];                   // qubit on column bit i (-1: none)
    int fj_pair[5][kMaxNbr], xk_pair[5][kMaxNbr];
    int n_col_bits;
    int n_pairs;
    int has_aj;
    // pair classes: 0 JJ, 1 KK, 2 JK, 3 JC, 4 KC, 5 CC ; i0/i1 local bit indices
    signed char cls[kMaxPairs], i0[kMaxPairs], i1[kMaxPairs];
};

struct __align__(16) KetDesc {
    const c128* src;                // read by pass 0
    c128* buf;                      // written by every storing pass, read by passes > 0
    const PassStep* steps;
    double* partial;                // [tiles] energy partials of the final pass (may be NULL)
    double sigma;                   // +1 / -1 : sign of the shift gate
    double escale;                  // multiplies the energy (1/(1+r^2) for scaled X-shifted kets)
    int n_pass;
    int shift_kind;                 // -1 none, 0 ZZ on (sb0, sb1), 1 X on sb0   (physical bits)
    int sb0, sb1;
    int cls;                        // pass type of its pass 0 (0 = L first, 1 = H first)
    int map_src;                    // tensor-map pair (L view, H view) of `src` in LaunchArgs::maps
    const c128* cross;              // linear mode: final state a = U phi; the energy pass also reduces Re <a|M|ket>
    double* partial2;               // [tiles] partials of that cross term
    double escale2;
    int map_buf;                    // tensor-map pair of `buf`
    int pad_;
};

static_assert(sizeof(KetDesc) % 16 == 0, "KetDesc is copied to shared memory in 16-byte pieces");

struct SetupJob {
    long long row_pre;              // row index into the angle table, -1 = none
    long long row_cur;              // -1 = final pass
    int type;
    int flags;
};

struct LaunchArgs {
    long long* trace;               // DQ_TRACE only: [item][warp][8] timestamps
    const KetDesc* kets;
    const CUtensorMap* maps;        // [2 * index + type]: TMA views of every state buffer the kets name
    const double2* tc;
    const double* mdiag;
    unsigned* counters;             // [0] next item, [1 + g] tiles done of ket g
    int n_kets;
    int group;                      // kets that run concurrently; ket k >= group reuses the work buffer of ket k - group and
                                    // starts when that one has finished (a launch CHAINS n_groups groups of `group` kets)
    int n_groups;
    int max_pass;
    unsigned long long inv_group, inv_pass;   // floor(2^64 / d) + 1: x / d = umul64hi(x, inv) for every 32-bit x (d >= 2)
    int tiles_log2;
    int sub_log2;                   // a work item is 2^sub_log2 consecutive tiles: one atomic / poll / release per item
    int ipp_log2;                   // items per (ket, pass) = tiles >> sub_log2
    int h_c1_shift;                 // H view: coordinate 1 of tile t is t << h_c1_shift
    int h_sw64;                     // H tiles land in the 64-byte swizzle (rows of 4 amplitudes, n = 20); see process_tile
    double r, ca, sa, c2a, s2a;     // shift gate: r, cos/sin(atan r), cos/sin(2 atan r)
    double rtau;                    // r / (1 + r^2): second coefficient of the shift gate's lifting butterfly
    const PassStep* steps0;         // first PassStep of the run: (ps - steps0) << n_col_bits is that pass-step's column table
    int n_col_bits;
    TypeGeom geom[2];
};

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
// Poll of a dependency counter on the hot path.  Relaxed on purpose: an acquire load holds back every later memory
// instruction of warp 0 for an L2 round trip (measured: ~800 cycles per tile, and the other warps wait for warp 0 at the
// next barrier).  What depends on the counter is only the TMA load of the tile, which is issued under a branch on the
// polled value, reads L2 directly (the producers' bulk stores are complete and fenced before they bump the counter)
// and never goes through this SM's L1; nothing else the kernel reads is written during the launch.
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Reserve the next work item (called by thread 0 of a team only).  ptxas turns an atomic add on a warp-uniform address
// into a warp-aggregated atomic whose result is broadcast with a shuffle right away, i.e. it waits for the L2 round trip
// (~800 cycles per tile, measured) that reserving one item ahead is meant to hide.  An address that formally depends on
// the lane id (lane 0 -> counters[0], the only lane that ever gets here) keeps it a plain atomic.
__device__ __forceinline__ unsigned take_item(unsigned* counters) {
    unsigned lane, v;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    asm volatile("atom.relaxed.gpu.global.add.u32 %0, [%1], 1;" : "=r"(v) : "l"(counters + lane) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int gather3(size_t x, const int* pos, const int* msk) {
    return (int)(((x >> pos[0]) & msk[0]) | (((x >> pos[1]) & msk[1]) << 1) | (((x >> pos[2]) & msk[2]) << 2));
}

// exp(-i theta X) on register bit B.  SCALED: rc = (1, tan) -> a' = a - i t b (cos folded elsewhere).
template <bool SCALED, int B>
__device__ __forceinline__ void rot_bit(c128 (&v)[kRegs], const double2 rc) {
    if (DQ_EXP & 1) return;
#pragma unroll
    for (int j = 0; j < kRegs; ++j) {
        if (j & (1 << B)) continue;
        const c128 a = v[j], b = v[j | (1 << B)];
        if (SCALED) {
            v[j] = make_double2(fma(rc.y, b.y, a.x), fma(-rc.y, b.x, a.y));
            v[j | (1 << B)] = make_double2(fma(rc.y, a.y, b.x), fma(-rc.y, a.x, b.y));
        } else {
            v[j] = make_double2(fma(rc.y, b.y, rc.x * a.x), fma(-rc.y, b.x, rc.x * a.y));
            v[j | (1 << B)] = make_double2(fma(rc.y, a.y, rc.x * b.x), fma(-rc.y, a.x, rc.x * b.y));
        }
    }
}

// LIFTING form of the same rotation, every FMA accumulating IN PLACE (destination = addend), rc = (t, tau) = (tan, sin cos):
//     a <- a - i t b            = a'/cos                      (the scaled butterfly for the bit-0 half)
//     b <- b - i tau a(new)     = (b - i t a)/(1 + t^2) = b' cos
// Same 4 DFMA per pair as the scaled form, but no value ever changes register -- the accumulator pattern of a GEMM inner
// loop -- so ptxas keeps the 128 amplitude registers in place across loop back-edges and ONE copy of the 320-DFMA run
// serves all four rotation rounds of a tile (the straight-line form needs four: 37 KB of code per pass type against a
// 32 KB instruction cache, the measured limiter of round 1).  The price is a per-amplitude pending factor
// (bit ? cos : 1/cos) for every rotated bit; it is a product over index bits, exactly the shape of the phase tables,
// and k_setup folds its inverse into them (see "resolve" there): no instruction is spent on it here.
template <int B>
__device__ __forceinline__ void lift_bit(c128 (&v)[kRegs], const double2 rc) {
#pragma unroll
    for (int j = 0; j < kRegs; ++j) {
        if (j & (1 << B)) continue;
        const int k = j | (1 << B);
        v[j].x = fma(rc.x, v[k].y, v[j].x);
        v[j].y = fma(-rc.x, v[k].x, v[j].y);
        v[k].x = fma(rc.y, v[j].y, v[k].x);
        v[k].y = fma(-rc.y, v[j].x, v[k].y);
    }
}
__device__ __forceinline__ void lift_run(c128 (&v)[kRegs], const double2* rc) {
    lift_bit<0>(v, rc[0]);
    lift_bit<1>(v, rc[1]);
    lift_bit<2>(v, rc[2]);
    lift_bit<3>(v, rc[3]);
    lift_bit<4>(v, rc[4]);
}

// Straight-line on purpose: a branch around a butterfly block makes ptxas reconcile the 128
// amplitude registers at the join with one move per FMA (measured: 2150 IMAD.MOV in v0).
// An inactive slot holds the identity (1, 0) and costs only its FMAs.
template <bool SCALED>
__device__ __forceinline__ void rot_run(c128 (&v)[kRegs], const double2* rc, const int override_bit,
                                        const double2 override_rc) {
    double2 r0 = rc[0], r1 = rc[1], r2 = rc[2], r3 = rc[3], r4 = rc[4];
    if (override_bit == 0) r0 = override_rc;
    if (override_bit == 1) r1 = override_rc;
    if (override_bit == 2) r2 = override_rc;
    if (override_bit == 3) r3 = override_rc;
    if (override_bit == 4) r4 = override_rc;
    rot_bit<SCALED, 0>(v, r0);
    rot_bit<SCALED, 1>(v, r1);
    rot_bit<SCALED, 2>(v, r2);
    rot_bit<SCALED, 3>(v, r3);
    rot_bit<SCALED, 4>(v, r4);
}

// Geometry of the two pass types.  A tile index has 12 bits t0..t11 (L: physical bits 0..11; H: `a` low
// spectator bits, then physical bits 10..n-1).  Shared memory holds the tile in natural order under the TMA
// 128-byte swizzle  idx ^ ((idx >> 3) & 7)  (16-byte units), which pairs t0|t3, t1|t4, t2|t5.  A register
// round is conflict-free when the three lowest lane bits take one bit of each pair and the other bit of every
// pair is constant across a quarter-warp (a register bit or lane bit 3 / 4):
//   TYPE 0 (L): K = {t3,t4,t5,t8,t9}, J = {t0,t1,t2,t6,t7}; the warp owns t10,t11 in both rounds, so the whole
//               pass never leaves the warp's 1024 amplitudes.
//   TYPE 1 (H): K = {t5..t9}, J = {t2,t3,t4,t10,t11}; spare bits t0,t1.  J_L and J_H share at most one
//               physical bit for every n: a ZZ shift gate always finds a pass type in which its pair is not J-J.
template <int TYPE> struct Geo;
template <> struct Geo<0> {
    static constexpr int spare_shift = 10;
    __host__ __device__ static constexpr int regK(int j) { return ((j & 7) << 3) | ((j >> 3) << 8); }
    __host__ __device__ static constexpr int regJ(int j) { return (j & 7) | ((j >> 3) << 6); }
    __device__ static __forceinline__ int baseK(int tid) { return (tid & 7) | (((tid >> 3) & 3) << 6) | ((tid >> 5) << 10); }
    __device__ static __forceinline__ int baseJ(int tid) { return ((tid & 7) << 3) | (((tid >> 3) & 3) << 8) | ((tid >> 5) << 10); }
    __device__ static __forceinline__ int kbits(int i) { return ((i >> 3) & 7) | (((i >> 8) & 3) << 3); }
    __device__ static __forceinline__ int kslot(int tb) { return (tb >= 3 && tb <= 5) ? tb - 3 : ((tb == 8 || tb == 9) ? tb - 5 : -1); }
    __device__ static __forceinline__ int jslot(int tb) { return (tb >= 0 && tb <= 2) ? tb : ((tb == 6 || tb == 7) ? tb - 3 : -1); }
};
template <> struct Geo<1> {
    static constexpr int spare_shift = 0;
    __host__ __device__ static constexpr int regK(int j) { return j << 5; }
    __host__ __device__ static constexpr int regJ(int j) { return ((j & 7) << 2) | ((j >> 3) << 10); }
    __device__ static __forceinline__ int baseK(int tid) { return (tid & 31) | ((tid >> 5) << 10); }
    __device__ static __forceinline__ int baseJ(int tid) { return (tid & 3) | (((tid >> 2) & 7) << 5) | ((tid >> 5) << 8); }
    __device__ static __forceinline__ int kbits(int i) { return (i >> 5) & 31; }
    __device__ static __forceinline__ int kslot(int tb) { return (tb >= 5 && tb <= 9) ? tb - 5 : -1; }
    __device__ static __forceinline__ int jslot(int tb) { return (tb >= 2 && tb <= 4) ? tb - 2 : ((tb == 10 || tb == 11) ? tb - 7 : -1); }
};
// host copies of the two bit lists (build_type, fused_j_sets)
static const int kJBits[2][5] = {{0, 1, 2, 6, 7}, {2, 3, 4, 10, 11}};
static const int kKBits[2][5] = {{3, 4, 5, 8, 9}, {5, 6, 7, 8, 9}};
__host__ __device__ __forceinline__ constexpr int swz(int i) { return i ^ ((i >> 3) & 7); }
// swz(thread part | register part) = swz(thread part) ^ swz(register part): the swizzle is linear over XOR and the two
// parts share no bit.  `st` is the thread's (per-tile opaque) swizzled base, `c` the compile-time register part; only
// the low three bits really XOR, the rest adds -- one LOP3 per distinct low part and an immediate offset per access.
__device__ __forceinline__ int slot(int st, int c) { return (st ^ (swz(c) & 7)) + (swz(c) & ~7); }

#if DQ_TRACE
#define TRACE(A, item, slot) do { if ((threadIdx.x & 31) == 0 && (A).trace) (A).trace[(size_t)(item) * 48 + ((threadIdx.x >> 5) & 3) * 8 + (slot)] = clock64(); } while (0)
// extras of an item (written by thread 0 of the team that loads it): 0 load issued at, 1 site (1, 2, 3; 9 = cold path), 2 pass type
#define TRACEX(A, item, k, val) do { if ((A).trace) (A).trace[(size_t)(item) * 48 + 32 + (k)] = (long long)(val); } while (0)
#else
#define TRACE(A, item, slot) do { } while (0)
#define TRACEX(A, item, k, val) do { } while (0)
#endif

// ---- asynchronous global -> shared copies: LDGSTS for the small tables, TMA for the tiles ----------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
    unsigned done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
// Team-local barriers (barrier 0 stays the CTA-wide __syncthreads of the prologue):
//   1 + team : the team's "__syncthreads";   3 + team : "this team is done with its tile buffer" -- warps 1..3
//   only announce it, warp 0 waits and then issues the store / releases the buffer.
__device__ __forceinline__ void team_sync(int team) { asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(kTeamThreads) : "memory"); }
__device__ __forceinline__ void tile_done_arrive(int team) { asm volatile("bar.arrive %0, %1;" ::"r"(3 + team), "n"(kTeamThreads) : "memory"); }
__device__ __forceinline__ void tile_done_sync(int team) { asm volatile("bar.sync %0, %1;" ::"r"(3 + team), "n"(kTeamThreads) : "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // stores have read their smem
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }         // stores are complete

// Ownership of the three tile buffers.  Only thread 0 of a team touches this.
struct CtaShared {
    unsigned long long full[kBufs]; // tile landed (TMA complete_tx)
    // One word, so that a claim is one load and one compare-and-swap:
    //   bits 0..2  buffers nobody owns          bits 3..5  the parity the next user of buffer b waits for on full[b]
    //   bit  6     the team that may claim a free buffer AHEAD of need (they alternate: without this the team that
    //              released a buffer takes it straight back and the other one never prefetches)
    //   bits 7..8  team t still has work
    unsigned state;
    // Pacing (see pace_team): how many of its tiles team t has taken past the J1 rotations / past the mid-tile barrier
    unsigned mid2[kTeams], mid3[kTeams];
};

// The tile body is ~48 KiB of straight-line code per pass type, the per-SM instruction cache holds 32 KiB, and four warps
// in step share one fetch -- so every line was fetched from the GPC-level cache once per team and tile, and that cache
// ran at 96 % of its request rate (ncu gcc__cache_requests_type_instruction): the kernel was instruction-fetch bound.
// Team 1 therefore FOLLOWS team 0 through the same code a few thousand cycles behind, close enough to hit the lines team 0
// has just brought in: it starts its k-th tile only once team 0 is past the J1 rotations of its k-th tile, and team 0
// starts its next tile only once team 1 is past the J1 rotations of its previous one.  The same distance is what lets
// the spare tile buffer alternate between the teams.  Waits are bounded: the other team may be spinning on a dependency
// that only our own pending release can satisfy.
__device__ __forceinline__ void pace_team(CtaShared& cs, int team, unsigned my_tiles) {
    if (!((*reinterpret_cast<volatile unsigned*>(&cs.state) >> (7 + (team ^ 1))) & 1u)) return;
#if DQ_PACE_TIGHT
    volatile unsigned* mark = team == 1 ? &cs.mid2[0] : &cs.mid3[1];
#else
    volatile unsigned* mark = &cs.mid2[team ^ 1];
#endif
    const unsigned need = team == 1 ? my_tiles + 1u : my_tiles;
    for (int spin = 0; spin < 200 && *mark < need; ++spin) {
        __nanosleep(20);
        if (!((*reinterpret_cast<volatile unsigned*>(&cs.state) >> (7 + (team ^ 1))) & 1u)) return;
    }
}
// Claim a free buffer: returns buffer | parity << 4, or -1.  `ahead`: a claim ahead of need honours the turn (unless the
// other team is gone) and passes it on.
__device__ __forceinline__ int claim_buffer(CtaShared& cs, int team, bool ahead) {
    unsigned st = *reinterpret_cast<volatile unsigned*>(&cs.state);
    for (;;) {
        if (ahead && ((st >> 6) & 1u) != (unsigned)team && ((st >> (7 + (team ^ 1))) & 1u)) return -1;
        const unsigned m = st & 7u;
        if (!m) return -1;
        const int b = __ffs(m) - 1;
        unsigned nw = (st & ~(1u << b)) ^ (8u << b);
        if (ahead) nw = (nw & ~64u) | ((unsigned)(team ^ 1) << 6);
        const unsigned old = atomicCAS(&cs.state, st, nw);
        if (old == st) return b | (int)(((st >> (3 + b)) & 1u) << 4);
        st = old;
    }
}
// Keep a buffer we already own for one more tile: only its parity advances.
__device__ __forceinline__ int take_parity(CtaShared& cs, int b) { return (int)((atomicXor(&cs.state, 8u << b) >> (3 + b)) & 1u); }
__device__ __forceinline__ void release_buffer(CtaShared& cs, int b) {
    __threadfence_block();
    atomicOr(&cs.state, 1u << b);
}

// One work item = (pass p, ket g, tile t_id); written to shared memory by thread 0.
struct ItemInfo {
    unsigned item;
    int p, g, t_id;                 // t_id = tile of THIS visit (an item is visited once per tile)
    int grp, sub, ip;               // item index inside its (ket, pass); tile index inside the item; item parity (sh.red slot)
    int valid;                      // p < n_pass of that ket
    int ready;                      // its dependency was satisfied when thread 0 polled AND its tile load is in flight
    int buf, par;                   // ... into this buffer, completing full[buf] with this parity
};

__device__ __forceinline__ void decode_item(const LaunchArgs& A, unsigned item, ItemInfo& I) {
    I.item = item;
    I.grp = (int)(item & ((1u << A.ipp_log2) - 1u));
    I.sub = 0;
    I.ip = 0;
    I.t_id = I.grp << A.sub_log2;
    const unsigned rest = item >> A.ipp_log2;          // group-major, then pass-major inside a group
    const unsigned r2 = A.group > 1 ? (unsigned)__umul64hi((unsigned long long)rest, A.inv_group) : rest;
    const unsigned gl = rest - r2 * (unsigned)A.group;
    const unsigned r3 = A.max_pass > 1 ? (unsigned)__umul64hi((unsigned long long)r2, A.inv_pass) : r2;
    I.p = (int)(r2 - r3 * (unsigned)A.max_pass);
    I.g = (int)(r3 * (unsigned)A.group + gl);
}

__device__ __forceinline__ bool item_valid(const LaunchArgs& A, const KetDesc* __restrict__ skets, const ItemInfo& I, unsigned total) {
    return I.item < total && I.g < A.n_kets && I.p < skets[I.g].n_pass;
}

// What an item waits for: pass p > 0 for every tile of pass p - 1 of its ket; pass 0 of a chained ket for the whole ket
// that used its work buffer before.  Returns false when there is nothing to wait for.
__device__ __forceinline__ bool item_dependency(const LaunchArgs& A, const KetDesc* __restrict__ skets, int p, int g,
                                                int& ctr, unsigned& need) {
    if (p > 0) {
        ctr = 1 + g;
        need = (unsigned)p << A.ipp_log2;
        return true;
    }
    if (g >= A.group) {
        ctr = 1 + g - A.group;
        need = (unsigned)skets[g - A.group].n_pass << A.ipp_log2;
        return true;
    }
    return false;
}

// One TMA instruction brings the whole 64 KiB tile (thread 0 only).  L view: {8 amplitudes, 256 rows, pairs of
// 2048 amplitudes}; H view: {2^w low amplitudes, low-column groups, high rows 10..17, high rows 18..} (make_maps).
__device__ __forceinline__ void issue_tile(const LaunchArgs& A, const KetDesc* __restrict__ kd, const int type,
                                           const int p, const int t_id, c128* __restrict__ tile,
                                           unsigned long long* __restrict__ full) {
    const CUtensorMap* map = A.maps + 2 * (p == 0 ? kd->map_src : kd->map_buf) + type;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(full, (unsigned)(sizeof(c128) * kTile));
    if (type == 0) {
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(tile)), "l"(map), "r"(smem_u32(full)), "r"(0), "r"(0), "r"(2 * t_id) : "memory");
    } else {
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(smem_u32(tile)), "l"(map), "r"(smem_u32(full)), "r"(0), "r"(t_id << A.h_c1_shift), "r"(0), "r"(0) : "memory");
    }
}

// The mirror image: one TMA instruction writes the finished tile back (thread 0, after the team's done-barrier).
__device__ __forceinline__ void store_tile(const LaunchArgs& A, const KetDesc* __restrict__ kd, const int type,
                                           const int t_id, const c128* __restrict__ tile) {
    const CUtensorMap* map = A.maps + 2 * kd->map_buf + type;
#if DQ_STORE_EVICT_LAST
    // the next pass reads this tile back: ask L2 to keep it (ncu: without the hint 81 % of the bulk-store bytes went to DRAM)
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    if (type == 0) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
                     ::"l"(map), "r"(smem_u32(tile)), "r"(0), "r"(0), "r"(2 * t_id), "l"(pol) : "memory");
    } else {
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%2, %3, %4, %5}], [%1], %6;"
                     ::"l"(map), "r"(smem_u32(tile)), "r"(0), "r"(t_id << A.h_c1_shift), "r"(0), "r"(0), "l"(pol) : "memory");
    }
#else
    if (type == 0) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                     ::"l"(map), "r"(smem_u32(tile)), "r"(0), "r"(0), "r"(2 * t_id) : "memory");
    } else {
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                     ::"l"(map), "r"(smem_u32(tile)), "r"(0), "r"(t_id << A.h_c1_shift), "r"(0), "r"(0) : "memory");
    }
#endif
    bulk_commit();
}

__device__ __forceinline__ void prefetch_tables(const PassStep* __restrict__ ps, PassStep* __restrict__ slot) {
    const char* s = reinterpret_cast<const char*>(ps);
    char* d = reinterpret_cast<char*>(slot);
    for (int i = threadIdx.x & (kTeamThreads - 1); i < (int)(sizeof(PassStep) / 16); i += kTeamThreads) cp_async16(d + 16 * i, s + 16 * i);
}

// Which PassStep sits in each of the two shared-memory table slots.  Two scalars, not an array: a dynamically
// indexed pointer array lives in local memory, and every LDL after a fence is an L2 round trip.
struct CachedSteps {
    const PassStep* s0 = nullptr;
    const PassStep* s1 = nullptr;
    __device__ __forceinline__ const PassStep* get(int i) const { return i ? s1 : s0; }
    __device__ __forceinline__ void set(int i, const PassStep* v) { if (i) s1 = v; else s0 = v; }
};

struct Shared {
    ItemInfo info[2];
    double red[2][kTeamThreads / 32];   // energy partials of the item in slot `cur` (consumed one item later)
    double red2[2][kTeamThreads / 32];  // cross-term partials (linear mode)
};

// Completion of an item is published one half-item late: thread 0 keeps the record here and releases it
// right after the next item's mid-tile barrier, when the bulk store it covers was issued thousands of cycles
// ago and both the wait and the fence return at once.  (Measured: a fence directly after the stores cost ~25%.)
struct Pending {
    double* partial;                // where to write the energy partial (NULL = none)
    double* partial2;               // cross-term partial (NULL = none)
    double escale, escale2;
    int g;                          // ket whose counter is bumped, -1 = nothing pending
    int slot;                       // sh.red slot
};

__device__ __forceinline__ void flush_pending(const LaunchArgs& A, Shared& sh, Pending& pd) {
    if (pd.g < 0) return;
    if (pd.partial) *pd.partial = (sh.red[pd.slot][0] + sh.red[pd.slot][1] + sh.red[pd.slot][2] + sh.red[pd.slot][3]) * pd.escale;
    if (pd.partial2) *pd.partial2 = (sh.red2[pd.slot][0] + sh.red2[pd.slot][1] + sh.red2[pd.slot][2] + sh.red2[pd.slot][3]) * pd.escale2;
    // The item's bulk-tensor store is complete (performed at L2) before the counter moves; the consumers read the tile with
    // a TMA load from L2.  No gpu-scope fence: it cost thread 0 ~500 cycles per tile and orders nothing that is still in flight.
    bulk_wait_all();
    atomicAdd(&A.counters[1 + pd.g], 1u);
    pd.g = -1;
}

#ifndef DQ_PROBE_ONLY
// Everything between "the tile is in shared memory" and "the tile is stored / reduced" for one pass type.
// `nxt_raw` is the raw index of the following item (valid in thread 0 only); thread 0 turns it into
// sh.info[nb] half-way through, and as soon as that item's dependency holds and a tile buffer is free it
// issues the TMA load of its tile -- usually a whole half-tile before this one is finished.
template <bool SCALED, bool AJ, bool CROSS, int TYPE>
__device__ __forceinline__ void process_tile(const LaunchArgs& A, const KetDesc* __restrict__ kd, const PassStep& P,
                                             c128* __restrict__ tile, Shared& sh, const KetDesc* __restrict__ skets,
                                             const ItemInfo& I,
                                             const unsigned nxt_raw, const int nb, const unsigned total,
                                             PassStep* __restrict__ cache, CachedSteps& cached_ps,
                                             const int cb, int& next_cb, bool& next_tables_new, Pending& pd,
                                             CtaShared& cs, c128* __restrict__ tiles, const int team, const int my_buf,
                                             int& my_free, const unsigned my_tiles) {
    using G = Geo<TYPE>;
    const TypeGeom& T = A.geom[TYPE];
    const int tid = threadIdx.x & (kTeamThreads - 1);

    const int flags = P.flags;
    const int p = I.p, t_id = I.t_id;
    const bool last_sub = I.sub + 1 == (1 << A.sub_log2);

    // tile geometry
    size_t tbase, xK, xJ;
    const int iK = G::baseK(tid);
    const int iJ = G::baseJ(tid);
    // Opaque per tile: otherwise the 2 x 8 swizzled slot bases are hoisted out of the item loop, spilled, and
    // re-read from local memory right after the dependency poll has invalidated L1 (an L2 round trip per tile).
    int sK = swz(iK), sJ = swz(iJ);
    asm volatile("" : "+r"(sK), "+r"(sJ));
    if (TYPE == 0) {
        tbase = (size_t)t_id << kTileBits;
        xK = tbase + iK;
        xJ = tbase + iJ;
    } else {
        tbase = ((size_t)(t_id & ((1 << T.tid_lo_bits) - 1)) << T.a) | ((size_t)(t_id >> T.tid_lo_bits) << T.high_end);
        xK = tbase + ((size_t)(iK & T.lowmask) | ((size_t)(iK >> T.a) << 10));
        xJ = tbase + ((size_t)(iJ & T.lowmask) | ((size_t)(iJ >> T.a) << 10));
    }

    // shift gate of the estimator (pass 0 only), expressed as data so the amplitude code stays straight-line
    const int shift_kind = (p == 0) ? kd->shift_kind : -1;
    const double sigma = kd->sigma;
    int ovK = -1, ovJ = -1;
    int tb0 = -1, tb1 = -1;                  // tile bits of the shift operands
    if (shift_kind >= 0) {
        const int s0 = kd->sb0, s1 = kd->sb1;
        if (TYPE == 0) { tb0 = s0 < kTileBits ? s0 : -1; tb1 = s1 < kTileBits ? s1 : -1; }
        else {
            tb0 = s0 < T.a ? s0 : (s0 >= 10 ? s0 - 10 + T.a : -1);
            tb1 = s1 < T.a ? s1 : (s1 >= 10 ? s1 - 10 + T.a : -1);
        }
        if (shift_kind == 1) {
            ovK = G::kslot(tb0);
            ovJ = G::jslot(tb0);
        }
    }
    // (I + i sigma r X) = exp(-i theta X) / cos(theta) with tan(theta) = -sigma r
    const double2 shift_rc = SCALED ? make_double2(1.0, -sigma * A.r) : make_double2(A.ca, -sigma * A.sa);

    // base-phase column entry: an L2 round trip, issued now and consumed after the J1 rotations
    const c128 phi_tc = __ldg(A.tc + P.tc_offset + ((unsigned)((iJ >> G::spare_shift) & 3) | ((unsigned)t_id << 2)));
    const unsigned trace_item = sh.info[nb ^ 1].item;
    (void)trace_item;
    TRACE(A, trace_item, 1);
    c128 v[kRegs];
    // ---- outer-A : own slots -> registers, K-bit rotations of the previous step -----------------------
#pragma unroll
    // thread 0: claim a free buffer for the next item and start its load (tried at several points of the tile)
    bool dep_ok = false;
    ItemInfo nI;
    nI.valid = 0;
    nI.ready = 0;
    nI.buf = 0;
    nI.par = 0;
    // Before the mid-tile barrier the claim is published with the item; after it (late = true) thread 0 patches the
    // published item -- every thread reads it behind the team barrier at the top of the loop.
    auto try_issue = [&](int site, bool late) {
        const int got = claim_buffer(cs, team, true);
        if (got < 0) return;
        const int b = got & 15;
        TRACEX(A, nI.item, 0, clock64());
        TRACEX(A, nI.item, 1, site);
        const int pr = got >> 4;
        nI.buf = b;
        nI.par = pr;
        nI.ready = 1;
        if (late) {
            sh.info[nb].buf = b;
            sh.info[nb].par = pr;
            sh.info[nb].ready = 1;
        }
        const KetDesc* __restrict__ nkd = skets + nI.g;
        issue_tile(A, nkd, (nI.p + nkd->cls) & 1, nI.p, nI.t_id, tiles + (size_t)b * kTile, &cs.full[b]);
    };
    if (TYPE == 0) {
#pragma unroll
        for (int j = 0; j < kRegs; ++j) v[j] = tile[slot(sK, G::regK(j))];
    } else {
        // An H tile of a 20-qubit state has rows of 4 amplitudes (64 B).  A TMA box row narrower than the swizzle span
        // is padded to the span, so those tiles land in the 64-byte swizzle  idx ^ ((idx >> 3) & 3)  instead; this first
        // read is conflict-free in either (quarter-warp = t0,t1,t2; t3,t4 are lane bits 3,4).  The register part
        // j << 5 meets the 128-byte pattern only through bit 5 (XOR 4 for odd j) and the 64-byte pattern not at all.
        const int b0 = A.h_sw64 ? (iK ^ ((iK >> 3) & 3)) : sK;
        const int b1 = A.h_sw64 ? b0 : (b0 ^ 4);
#pragma unroll
        for (int j = 0; j < kRegs; ++j) v[j] = tile[((j & 1) ? b1 : b0) + G::regK(j)];
    }
    rot_run<SCALED>(v, P.rot[0], ovK, shift_rc);
    // from here on the tile lives in the 128-byte pattern; the 8 threads that share a 128-byte line are one quarter-warp
    if (TYPE == 1) __syncwarp();
#pragma unroll
    for (int j = 0; j < kRegs; ++j) if (!(DQ_EXP & 2)) tile[slot(sK, G::regK(j))] = v[j];
    // thread 0: the following item and its dependency counter
    unsigned polled = 0, need_next = 0;
    if (tid == 0) {
        if (my_free >= 0) {                  // the buffer of the previous tile becomes everybody's once its bulk store has read it
            bulk_wait_read();
            release_buffer(cs, my_free);
            my_free = -1;
        }
        if (!last_sub) {                     // next visit = next tile of the same item: nothing to fetch or poll
            nI = I;
            nI.sub = I.sub + 1;
            nI.t_id = I.t_id + 1;
            nI.valid = 1;
        } else {
            decode_item(A, nxt_raw, nI);
            nI.ip = I.ip ^ 1;
            nI.valid = item_valid(A, skets, nI, total);
            int ctr;
            if (nI.valid && item_dependency(A, skets, nI.p, nI.g, ctr, need_next))
                polled = ld_relaxed(&A.counters[ctr]);                         // consumed after the J1 rotations
        }
        nI.ready = 0;
        nI.buf = 0;
        nI.par = 0;
#if DQ_PACE_EARLY
        *reinterpret_cast<volatile unsigned*>(&cs.mid2[team]) = my_tiles;
#endif
    }
    TRACE(A, trace_item, 2);
    if (TYPE == 0) __syncwarp(); else team_sync(team);    // L: the exchange never leaves the warp's 1024 amplitudes

    // ---- inner : J-bit rotations, phase, J-bit rotations ---------------------------------------------
#pragma unroll
    for (int j = 0; j < kRegs; ++j) if (!(DQ_EXP & 2)) v[j] = tile[slot(sJ, G::regJ(j))];
    rot_run<SCALED>(v, P.rot[1], ovJ, shift_rc);
    if (tid == 0) {
        *reinterpret_cast<volatile unsigned*>(&cs.mid2[team]) = my_tiles;
        dep_ok = nI.valid && polled >= need_next;      // need_next = 0: nothing to wait for (or the next tile of this item)
        if (dep_ok) try_issue(2, false);
    }
    {
        const int kb = G::kbits(iJ);
        c128 phi = cmul(phi_tc, P.tkk[kb]);
#pragma unroll
        for (int m = 0; m < 5; ++m) {
            if (T.xk_msk[m][0]) {            // launch-uniform
                c128 w = P.xk[m][gather3(xJ, T.xk_pos[m], T.xk_msk[m])];
                if ((kb >> m) & 1) w.y = -w.y;
                phi = cmul(phi, w);
            }
        }
        c128 F[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) F[k] = P.fj[k][gather3(xJ, T.fj_pos[k], T.fj_msk[k])];
        if (shift_kind == 0) {               // ZZ shift gate exp(i sigma alpha z0 z1); never both operands in J
            const int j0b = G::jslot(tb0);
            const int j1b = G::jslot(tb1);
            const double z0 = ((xJ >> kd->sb0) & 1) ? -1.0 : 1.0;
            const double z1 = ((xJ >> kd->sb1) & 1) ? -1.0 : 1.0;
            // with the J operand at 0 (z = +1) the factor is exp(i sigma alpha z_other)
            phi = cmul(phi, make_double2(A.ca, sigma * A.sa * z0 * z1));
            const int jb = j0b >= 0 ? j0b : j1b;
            if (jb >= 0) {                   // flipping that J bit multiplies by exp(-2 i sigma alpha z_other)
                const double zo = j0b >= 0 ? z1 : z0;
                const c128 f = make_double2(A.c2a, -sigma * A.s2a * zo);
#pragma unroll
                for (int k = 0; k < 5; ++k)
                    if (k == jb) F[k] = cmul(F[k], f);
            }
        }
        if (!(DQ_EXP & 1))
#pragma unroll
        for (int b4 = 0; b4 < 2; ++b4) {
            const c128 p4 = b4 ? cmul(phi, F[4]) : phi;
#pragma unroll
            for (int b3 = 0; b3 < 2; ++b3) {
                const c128 p3 = b3 ? cmul(p4, F[3]) : p4;
#pragma unroll
                for (int b2 = 0; b2 < 2; ++b2) {
                    const c128 p2 = b2 ? cmul(p3, F[2]) : p3;
#pragma unroll
                    for (int b1 = 0; b1 < 2; ++b1) {
                        const c128 p1 = b1 ? cmul(p2, F[1]) : p2;
                        const int j = (b4 << 4) | (b3 << 3) | (b2 << 2) | (b1 << 1);
                        v[j] = cmul(v[j], p1);
                        v[j | 1] = cmul(v[j | 1], cmul(p1, F[0]));
                    }
                }
            }
        }
        if (DQ_EXP & 1) v[0] = cmul(v[0], cmul(phi, cmul(F[0], cmul(F[1], cmul(F[2], cmul(F[3], F[4]))))));
        if (AJ && !(DQ_EXP & 1)) {
#pragma unroll
            for (int j = 0; j < kRegs; ++j) v[j] = cmul(v[j], P.aj[j]);
        }
    }
    rot_run<SCALED>(v, P.rot[2], -1, shift_rc);
#pragma unroll
    for (int j = 0; j < kRegs; ++j) if (!(DQ_EXP & 2)) tile[slot(sJ, G::regJ(j))] = v[j];
    if (tid == 0) {                          // publish the following item
        if (dep_ok && !nI.ready) try_issue(3, false);
        sh.info[nb] = nI;
        *reinterpret_cast<volatile unsigned*>(&cs.mid3[team]) = my_tiles;
    }
    TRACE(A, trace_item, 3);
    team_sync(team);
    TRACE(A, trace_item, 4);
    if (tid == 0) flush_pending(A, sh, pd);  // previous item: all of its stores were ordered by this barrier

    // ---- outer-B : K-bit rotations of the new step; prefetch of the next tile; store or reduce --------
#pragma unroll
    for (int j = 0; j < kRegs; ++j) if (!(DQ_EXP & 2)) v[j] = tile[slot(sK, G::regK(j))];
    if (flags & F_ENERGY) {                  // energy pass: pull the observable (and the cross state) towards L1 now
        const double* __restrict__ mdp = A.mdiag + xK;
#pragma unroll
        for (int j = 0; j < kRegs; ++j) {
            const size_t o = TYPE == 0 ? (size_t)G::regK(j) : (size_t)T.offK[j];
            asm volatile("prefetch.global.L1 [%0];" ::"l"(mdp + o));
            if (CROSS && kd->cross) asm volatile("prefetch.global.L2 [%0];" ::"l"(kd->cross + xK + o));
        }
    }
    rot_bit<SCALED, 0>(v, P.rot[3][0]);
    {
        const ItemInfo& N = sh.info[nb];
        next_cb = cb;
        next_tables_new = false;
        if (N.valid) {                       // the tables of the next item, whether or not its tile is on its way yet
            const KetDesc* __restrict__ nkd = skets + N.g;
            const PassStep* nps = nkd->steps + N.p;
            if (nps != cached_ps.get(cb)) {
                next_cb = cb ^ 1;
                if (nps != cached_ps.get(next_cb)) {
                    prefetch_tables(nps, cache + next_cb);
                    cached_ps.set(next_cb, nps);
                    next_tables_new = true;
                }
            }
            cp_async_commit();
        }
    }
    TRACE(A, trace_item, 5);
    if (tid == 0 && dep_ok && !nI.ready) try_issue(4, true);
    rot_bit<SCALED, 1>(v, P.rot[3][1]);
    rot_bit<SCALED, 2>(v, P.rot[3][2]);
    rot_bit<SCALED, 3>(v, P.rot[3][3]);
    if (tid == 0 && dep_ok && !nI.ready) try_issue(5, true);
    TRACE(A, trace_item, 6);
    rot_bit<SCALED, 4>(v, P.rot[3][4]);
    {
        const bool do_store = (flags & F_STORE) != 0, do_energy = (flags & F_ENERGY) != 0;
        const double* __restrict__ md = A.mdiag + xK;
        const c128* __restrict__ cr = (CROSS && do_energy && kd->cross) ? kd->cross + xK : nullptr;
        double e = 0.0, e2 = 0.0;
        if (do_store) {                      // back into the landing layout of this buffer; one bulk store takes it from there
            if (TYPE == 0) {
#pragma unroll
                for (int j = 0; j < kRegs; ++j) tile[slot(sK, G::regK(j))] = v[j];
            } else {
                __syncwarp();                // 64-byte pattern: these are not the slots this thread has just read
                const int b0 = A.h_sw64 ? (iK ^ ((iK >> 3) & 3)) : sK;
                const int b1 = A.h_sw64 ? b0 : (b0 ^ 4);
#pragma unroll
                for (int j = 0; j < kRegs; ++j) tile[((j & 1) ? b1 : b0) + G::regK(j)] = v[j];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        if (tid < 32) tile_done_sync(team); else tile_done_arrive(team);
        if (tid == 0) {
            if (do_store) {
                store_tile(A, kd, TYPE, t_id, tile);
                my_free = my_buf;            // released at the top of the next tile, when the store has read it
            } else {
                release_buffer(cs, my_buf);
            }
            if (dep_ok && !nI.ready) try_issue(6, true);
            if (dep_ok && !nI.ready && do_store) {
                // nothing free and the next item is ready to go: take our own buffer back as soon as the store has read it
                // (the other warps of the team are on their way to the top of the loop and wait there)
                bulk_wait_read();
                my_free = -1;
                sh.info[nb].buf = my_buf;
                sh.info[nb].par = take_parity(cs, my_buf);
                sh.info[nb].ready = 1;
                TRACEX(A, nI.item, 0, clock64());
                TRACEX(A, nI.item, 1, 7);
                const KetDesc* __restrict__ nkd = skets + nI.g;
                issue_tile(A, nkd, (nI.p + nkd->cls) & 1, nI.p, nI.t_id, tile, &cs.full[my_buf]);
            }
        }
        if (do_energy) {
#pragma unroll
            for (int j = 0; j < kRegs; ++j) {
                const size_t o = TYPE == 0 ? (size_t)G::regK(j) : (size_t)T.offK[j];
                const double m = __ldg(md + o);
                e = fma(m, fma(v[j].x, v[j].x, v[j].y * v[j].y), e);
                if (CROSS && cr) {           // Re conj(a) ket
                    const c128 a0 = __ldcg(cr + o);
                    e2 = fma(m, fma(a0.x, v[j].x, a0.y * v[j].y), e2);
                }
            }
        }
        if (do_energy) {
            for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
            if ((tid & 31) == 0) sh.red[I.ip][tid >> 5] = I.sub == 0 ? e : sh.red[I.ip][tid >> 5] + e;
            if (CROSS && cr) {
                for (int o = 16; o > 0; o >>= 1) e2 += __shfl_xor_sync(0xffffffffu, e2, o);
                if ((tid & 31) == 0) sh.red2[I.ip][tid >> 5] = I.sub == 0 ? e2 : sh.red2[I.ip][tid >> 5] + e2;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// the persistent pass kernel
// ------------------------------------------------------------------------------------------
template <bool SCALED, bool AJ, bool CROSS>
__global__ void __launch_bounds__(kThreads, DQ_CTAS_PER_SM) k_fused_passes(const __grid_constant__ LaunchArgs A) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    // the TMA swizzle is a function of the shared-memory address: the tile buffers start on a 1 KiB boundary
    unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    c128* tiles = reinterpret_cast<c128*>(smem_raw);                                                  // kBufs buffers
    PassStep* cache_all = reinterpret_cast<PassStep*>(smem_raw + sizeof(c128) * kTile * kBufs);      // two table slots per team
    // Ket descriptors live in shared memory: the acquire / release fences of the item protocol invalidate
    // L1 (CCTL.IVALL), so every kd-> field read from global memory was an exposed L2 round trip per tile.
    KetDesc* skets = reinterpret_cast<KetDesc*>(smem_raw + sizeof(c128) * kTile * kBufs + 2 * kTeams * sizeof(PassStep));
    __shared__ Shared sh_all[kTeams];
    __shared__ __align__(8) CtaShared cs;

    const int team = threadIdx.x / kTeamThreads;
    const int tid = threadIdx.x & (kTeamThreads - 1);
    Shared& sh = sh_all[team];
    PassStep* cache = cache_all + 2 * team;
    if (threadIdx.x == 0) {
        for (int b = 0; b < kBufs; ++b) mbar_init(&cs.full[b], 1);
        cs.state = ((1u << kBufs) - 1u) | (((1u << kTeams) - 1u) << 7);
        for (int t = 0; t < kTeams; ++t) cs.mid2[t] = cs.mid3[t] = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const unsigned total = ((unsigned)A.max_pass * (unsigned)A.group * (unsigned)A.n_groups) << A.ipp_log2;
    const int nsub = 1 << A.sub_log2;
    CachedSteps cached_ps;
    int cur = 0, cb = 0;
    bool tables_new = false;
    int my_free = -1;                          // thread 0: the buffer its last bulk store may still be reading
    unsigned my_tiles = 0;                     // tiles this team has started
    Pending pd;
    pd.g = -1;
    pd.partial = nullptr;
    pd.partial2 = nullptr;
    pd.escale = 0.0;
    pd.escale2 = 0.0;
    pd.slot = 0;

    // thread 0 of a team reserves items ONE AHEAD: q_next is the raw index of the item after sh.info[cur], taken
    // from the global counter a whole tile earlier, so the atomic's round trip is never waited for
    unsigned q_next = 0;
    if (tid == 0) {
        const unsigned first = take_item(A.counters);
        q_next = take_item(A.counters);
        ItemInfo I;
        decode_item(A, first, I);
        I.valid = 0;                           // completed below, once the descriptors are in shared memory
        I.ready = 0;
        I.buf = 0;
        I.par = 0;
        sh.info[0] = I;
    }
    {
        const int4* src = reinterpret_cast<const int4*>(A.kets);
        int4* dst = reinterpret_cast<int4*>(skets);
        for (int i = threadIdx.x; i < A.n_kets * (int)(sizeof(KetDesc) / 16); i += kThreads) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    if (tid == 0) sh.info[0].valid = item_valid(A, skets, sh.info[0], total);
    __syncthreads();

    for (;;) {
        if (tid == 0) pace_team(cs, team, my_tiles);
        team_sync(team);                       // also: thread 0's late claim of this item's tile is visible to the team
        const ItemInfo I = sh.info[cur];
        if (I.item >= total) break;
        if (!I.valid) {                        // ragged group: this ket has no such pass; fetch another item
            team_sync(team);
            if (tid == 0) {
                flush_pending(A, sh, pd);
                ItemInfo N;
                decode_item(A, q_next, N);
                q_next = take_item(A.counters);
                N.valid = item_valid(A, skets, N, total);
                N.ready = 0;
                N.buf = 0;
                N.par = 0;
                sh.info[cur ^ 1] = N;
            }
            team_sync(team);
            cur ^= 1;
            continue;
        }
        const KetDesc* __restrict__ kd = skets + I.g;
        int buf = I.buf, par = I.par;
        ++my_tiles;
        if (tid == 0) TRACEX(A, I.item, 9, clock64());
        if (!I.ready) {                        // cold path: wait for the dependency, then fetch tile and tables
            if (tid == 0) {
                int got = claim_buffer(cs, team, false);   // a free buffer, else our own last one once its store has read it
                if (got < 0 && my_free >= 0) {
                    bulk_wait_read();
                    got = my_free | (take_parity(cs, my_free) << 4);
                    my_free = -1;
                }
                while (got < 0) {
                    __nanosleep(64);
                    got = claim_buffer(cs, team, false);
                }
                const int b = got & 15;
                int ctr;
                unsigned need;
                if (item_dependency(A, skets, I.p, I.g, ctr, need)) {
                    if (ld_acquire(&A.counters[ctr]) < need) {
                        flush_pending(A, sh, pd);  // always before spinning: the dependency may be our own tile
                        while (ld_acquire(&A.counters[ctr]) < need) __nanosleep(32);
                    }
                }
                const int pr = got >> 4;
                sh.info[cur].buf = b;
                sh.info[cur].par = pr;
                TRACEX(A, I.item, 0, clock64());
                TRACEX(A, I.item, 1, 9);
                issue_tile(A, kd, (I.p + kd->cls) & 1, I.p, I.t_id, tiles + (size_t)b * kTile, &cs.full[b]);
            }
            team_sync(team);
            buf = sh.info[cur].buf;
            par = sh.info[cur].par;
            const PassStep* ps = kd->steps + I.p;
            if (ps != cached_ps.get(cb)) {
                cb ^= 1;
                if (ps != cached_ps.get(cb)) {
                    prefetch_tables(ps, cache + cb);
                    cached_ps.set(cb, ps);
                    tables_new = true;
                }
            }
            cp_async_commit();
        }
        if (tid == 0) TRACEX(A, I.item, 10, clock64());
        TRACE(A, I.item, 0);
        cp_async_wait_all();                   // this thread's share of the tables landed
        if (tables_new) team_sync(team);       // tables in cache[cb] become visible to every thread of the team
        if (tid == 0) TRACEX(A, I.item, 3, clock64());
        mbar_wait(&cs.full[buf], (unsigned)par);   // the tile landed (TMA complete_tx)
        if (tid == 0) TRACEX(A, I.item, 4, clock64());
        c128* tile = tiles + (size_t)buf * kTile;
        unsigned nxt_raw = 0;
        if (tid == 0 && I.sub + 1 == nsub) {   // consumed inside process_tile; after the waits above, which would also wait for it
            nxt_raw = q_next;
            q_next = take_item(A.counters);
        }
        int next_cb = cb;
        bool next_tables_new = false;
        const PassStep& P = cache[cb];
        const int flags = P.flags;
        if (tid == 0) { TRACEX(A, I.item, 2, P.type); TRACEX(A, I.item, 6, blockIdx.x * kTeams + team + 1); TRACEX(A, I.item, 7, I.p); }
        if (P.type == 0)
            process_tile<false, AJ, CROSS, 0>(A, kd, P, tile, sh, skets, I, nxt_raw, cur ^ 1, total, cache, cached_ps, cb,
                                        next_cb, next_tables_new, pd, cs, tiles, team, buf, my_free, my_tiles);
        else
            process_tile<false, AJ, CROSS, 1>(A, kd, P, tile, sh, skets, I, nxt_raw, cur ^ 1, total, cache, cached_ps, cb,
                                        next_cb, next_tables_new, pd, cs, tiles, team, buf, my_free, my_tiles);
        TRACE(A, I.item, 7);
        if (tid == 0 && I.sub + 1 == nsub) {   // item complete: published after the next item's mid-tile barrier
            pd.g = I.g;
            pd.partial = ((flags & F_ENERGY) && kd->partial) ? kd->partial + I.grp : nullptr;
            pd.partial2 = (CROSS && (flags & F_ENERGY) && kd->cross) ? kd->partial2 + I.grp : nullptr;
            pd.escale = kd->escale;
            pd.escale2 = kd->escale2;
            pd.slot = I.ip;
        }
        cur ^= 1;
        cb = next_cb;
        tables_new = next_tables_new;
        if (tid == 0 && sh.info[cur].item < total) TRACEX(A, sh.info[cur].item, 8, clock64());
    }
    cp_async_wait_all();
    team_sync(team);
    if (tid == 0) {
        atomicAnd(&cs.state, ~(128u << team));
        bulk_wait_all();
        flush_pending(A, sh, pd);
    }
}

#endif  // DQ_PROBE_ONLY

// ------------------------------------------------------------------------------------------
// the persistent pass kernel, WARP-SPECIALISED LOOP form (scaled mode: |angle| <= 1)
// ------------------------------------------------------------------------------------------
// Same passes, tiles, TMA views and inter-CTA dependency counters as k_fused_passes.  Two things change.
//
// (1) The tile body is a loop.  The four rotation rounds {KA, J1, J2, KB} share ONE copy of the 320-DFMA butterfly run
//     (lifting form, lift_bit: every FMA accumulates in place, so the 128 amplitude registers stay put across the
//     back-edge); what differs between rounds -- the shared-memory slots read and written, the phase -- sits in uniform
//     branches in front of it.  ~900 instructions per pass type instead of 2330: both types fit the 32 KB instruction
//     cache (the straight-line body streamed 37 KB per tile from the GPC-level cache, round 1's measured limiter).
//
// (2) The protocol leaves the compute warps.  A CTA is three warpgroups:
//       WG0, WG1  consumer teams (4 warps, 128 threads, 232 registers each after setmaxnreg): wait for a full tile buffer,
//                 run the tile body, write the result back into the buffer, arrive on its `done` barrier.  Team t takes
//                 tiles k = t, t+2, ... of this CTA, tile k lives in buffer k mod 3.  No item decoding, no polling,
//                 no TMA, no atomics in these warps.
//       WG2       warp 8 = LOADER: reserves work items from the global counter (one ahead), skips items of ragged groups,
//                 waits for the buffer to be free and for the item's dependency (all tiles of the previous pass of its ket),
//                 then issues three async copies onto the buffer's `full` mbarrier: the tile (one TMA tensor load), the
//                 pass-step tables (2.6 KB bulk copy) and the four base-phase column entries of the tile (64 B bulk copy).
//                 warp 9 = STORER: waits for `done`, issues the TMA tensor store, frees the buffer when the store has read
//                 it, writes the tile's energy partial, and publishes the tile (per-ket counter) once the store is complete
//                 at L2 -- one tile late, so it never waits for a store it has just issued.
//                 (40 registers after setmaxnreg; warps 10, 11 only give their registers back.)
//     Every wait of the data path is an mbarrier; the only spinning on global memory is the loader's dependency poll.
struct WsSlot {                     // one per tile buffer; written by the loader before `full` completes
    int p, g, t_id, grp;
    int stop;                       // 1: no more tiles for the team that receives this buffer
    unsigned item;                  // raw work-item index (DQ_TRACE builds index the trace with it)
    unsigned tile_idx;              // the tile's part of the ten phase-table indices (see ws_pack_idx), formed by the loader
    int pad_;
    double2 tc[4];                  // base-phase column entries of this tile (bulk-copied from the column table)
};
struct WsShared {
    unsigned long long full[kBufs], done[kBufs], empty[kBufs];
    WsSlot slot[kBufs];
    double red[kBufs][kTeamThreads / 32];
    double red2[kBufs][kTeamThreads / 32];
    unsigned math_lock[kTeamThreads / 32];      // see math_acquire
    unsigned thread_idx[2][kTeamThreads];       // per pass type: a team thread's part of the ten phase-table indices
};

// The phase of a step looks up ten small tables (fj[5], xk[5]) with 3-bit indices gathered from up to three physical bits of the
// amplitude index xJ = tile part | thread part.  Bit gathering is linear over OR and the two parts share no bit, so each part is
// gathered once -- the thread part per launch (ws.thread_idx), the tile part per tile by the loader (WsSlot::tile_idx) -- and
// packed 3 bits per table: fj[k] at bit 3k, xk[m] at bit 15 + 3m.  The consumers then form an index with one shift and one
// AND instead of ~20 instructions and six constant-bank loads (the ten gathers were 4 % of the kernel's instructions, all of
// them on the latency chain at the head of the phase section).
__device__ __forceinline__ unsigned ws_pack_idx(const TypeGeom& T, const size_t x) {
    unsigned packed = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        packed |= (unsigned)gather3(x, T.fj_pos[k], T.fj_msk[k]) << (3 * k);
        packed |= (unsigned)gather3(x, T.xk_pos[k], T.xk_msk[k]) << (15 + 3 * k);
    }
    return packed;
}
constexpr int kWsThreads = 384;
constexpr int kWsConsumerRegs = 232, kWsProducerRegs = 40;     // 2 x 128 x 232 + 128 x 40 = 64512 <= 65536

__device__ __forceinline__ bool mbar_test(unsigned long long* bar, unsigned parity) {
    unsigned done;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_group() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// Warp i of team 0 and warp i of team 1 sit on the same SM sub-partition and share its FP64 pipe (one DFMA per two
// cycles: a single warp saturates it) and, with the other six warps, the shared-memory pipe.  Left alone, the two fall
// into a convoy -- both in their butterfly runs at half speed each, then both in their shared-memory exchanges -- and
// neither pipe is ever busy while the other is: measured 44 % FP64 and 47 % LSU, i.e. no overlap at all.  The lock below
// makes the math sections of the two warps mutually exclusive: whoever gets there second waits, runs alone at full
// speed afterwards, and from then on is in its exchange while the other one computes.
#ifndef DQ_MATH_LOCK
#define DQ_MATH_LOCK 0
#endif
__device__ __forceinline__ void math_acquire(WsShared& ws, int w) {
    if ((threadIdx.x & 31) == 0)
        while (atomicCAS(&ws.math_lock[w], 0u, 1u) != 0u) __nanosleep(32);
    __syncwarp();
}
__device__ __forceinline__ void math_release(WsShared& ws, int w) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) atomicExch(&ws.math_lock[w], 0u);
}

// The tile body of a consumer team: everything between "the buffer is full" and "the result is back in the buffer".
// ONE body serves both pass types: the type is a run-time value, and only the small blocks that move amplitudes between
// registers and shared memory are written out per type (their slot offsets are compile-time constants of Geo<>); the
// butterfly run, the phase and the energy reduction exist once.  (Two fully inlined per-type bodies behind one branch
// made ptxas spill; this form does not, and it is half the code.)
#define DQ_FOR_REGS(stmt) _Pragma("unroll") for (int j = 0; j < kRegs; ++j) { stmt; }
#if DQ_TRACE
#define WS_TRACE(A, item, slot) do { if ((threadIdx.x & 31) == 0 && (A).trace) (A).trace[(size_t)(item) * 48 + ((threadIdx.x >> 5) & 3) * 8 + (slot)] = clock64(); } while (0)
#else
#define WS_TRACE(A, item, slot) do { } while (0)
#endif
template <bool AJ, bool CROSS>
__device__ __forceinline__ void ws_tile(const LaunchArgs& A, WsShared& ws, const KetDesc* __restrict__ skets,
                                        PassStep& P, c128* __restrict__ tile, const int b, const int team) {
    using G0 = Geo<0>;
    using G1 = Geo<1>;
    const int type = P.type;
    const TypeGeom& T = A.geom[type];
    int tid = threadIdx.x & (kTeamThreads - 1);
    const WsSlot& S = ws.slot[b];

    auto tile_bit_of = [&](int s) { return s < T.a ? s : (s >= T.start ? s - T.start + T.a : -1); };   // physical bit -> tile bit

    // Pass 0 of a shifted ket: the X shift gate (I + i sigma r X) = exp(-i theta X)/cos(theta), tan(theta) = -sigma r, rides on
    // the KA or J1 round as one more butterfly.  This buffer's copy of the pass-step tables is private to the tile, so the
    // gate is PATCHED into it -- the lifting pair (t, t/(1+t^2)) into the (otherwise identity) rotation slot, and the factor
    // (1 + r^2) that the lifting form leaves its bit-1 half short of (relative to the scaled form `escale` expects) into the
    // phase tables -- and the round loop below stays free of shift-gate code.
    if (S.p == 0) {
        const KetDesc* __restrict__ kd = skets + S.g;
        if (kd->shift_kind == 1) {
            const int tb0 = tile_bit_of(kd->sb0);
            const int ovK = type == 0 ? G0::kslot(tb0) : G1::kslot(tb0), ovJ = type == 0 ? G0::jslot(tb0) : G1::jslot(tb0);
            const double r2p1 = fma(A.r, A.r, 1.0);
            if (tid == 0) {
                const double2 rc = make_double2(-kd->sigma * A.r, -kd->sigma * A.rtau);
                if (ovK >= 0) P.rot[0][ovK] = rc;
                if (ovJ >= 0) P.rot[1][ovJ] = rc;
            }
            if (ovK >= 0 && tid < 32 && ((tid >> ovK) & 1)) { P.tkk[tid].x *= r2p1; P.tkk[tid].y *= r2p1; }
            if (ovJ >= 0 && tid < 8) { P.fj[ovJ][tid].x *= r2p1; P.fj[ovJ][tid].y *= r2p1; }
        }
        team_sync(team);
    }

    c128 v[kRegs];
#if DQ_WS_PIPE
    // Software-pipelined form of the round loop.  Each round is  [head: amplitudes into registers, or the phase] [butterfly
    // stages 0..3, one shared copy] [tail: stage 4, its stores issued pair by pair as the results appear].  The stores of an
    // exchange ride on the last butterfly stage of the run that produced the data instead of following it: the LSU drains them
    // while the FP64 pipe is still busy (with two consumer warps per scheduler nothing else would cover them).
#define DQ_LIFT4_ST(ADDR)                                                                         \
    _Pragma("unroll") for (int j = 0; j < 16; ++j) {                                              \
        const int k = j | 16;                                                                     \
        v[j].x = fma(rc4.x, v[k].y, v[j].x);                                                      \
        v[j].y = fma(-rc4.x, v[k].x, v[j].y);                                                     \
        v[k].x = fma(rc4.y, v[j].y, v[k].x);                                                      \
        v[k].y = fma(-rc4.y, v[j].x, v[k].y);                                                     \
        tile[ADDR(j)] = v[j];                                                                     \
        tile[ADDR(k)] = v[k];                                                                     \
    }
#define DQ_A_K0(j) slot(sK, G0::regK(j))
#define DQ_A_K1(j) slot(sK, G1::regK(j))
#define DQ_A_J0(j) slot(sJ, G0::regJ(j))
#define DQ_A_J1(j) slot(sJ, G1::regJ(j))
#define DQ_A_LAND1(j) ((((j) & 1) ? b1 : b0) + G1::regK(j))
    const int flags_all = P.flags;
#pragma unroll 1
    for (int r = 0; r < 4; ++r) {
        asm volatile("" : "+r"(tid));
        const int iK = type == 0 ? G0::baseK(tid) : G1::baseK(tid);
        const int iJ = type == 0 ? G0::baseJ(tid) : G1::baseJ(tid);
        const int sK = swz(iK), sJ = swz(iJ);
        const int b0 = A.h_sw64 ? (iK ^ ((iK >> 3) & 3)) : sK;       // H landing layout: 64-byte swizzle at n = 20
        const int b1 = A.h_sw64 ? b0 : (b0 ^ 4);
        // ---- head ------------------------------------------------------------------------------------------------
        if (r == 0) {
            if (type == 0) {
                DQ_FOR_REGS(v[j] = tile[DQ_A_K0(j)])
            } else {
                DQ_FOR_REGS(v[j] = tile[DQ_A_LAND1(j)])
            }
        } else if (r == 1) {
            if (type == 0) {
                DQ_FOR_REGS(v[j] = tile[DQ_A_J0(j)])
            } else {
                DQ_FOR_REGS(v[j] = tile[DQ_A_J1(j)])
            }
        } else if (r == 2) {
            // ---- phase of the new step (the J1 rotations are done) ------------------------------------------
            const unsigned idx = ws.thread_idx[type][tid] | S.tile_idx;          // ten 3-bit table indices (ws_pack_idx)
            const int kb = type == 0 ? G0::kbits(iJ) : G1::kbits(iJ);
            c128 phi = cmul(S.tc[(iJ >> (type == 0 ? G0::spare_shift : G1::spare_shift)) & 3], P.tkk[kb]);
#pragma unroll
            for (int m = 0; m < 5; ++m) {
                if (T.xk_msk[m][0]) {            // launch-uniform
                    c128 w = P.xk[m][(idx >> (15 + 3 * m)) & 7u];
                    if ((kb >> m) & 1) w.y = -w.y;
                    phi = cmul(phi, w);
                }
            }
            c128 F[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) F[k] = P.fj[k][(idx >> (3 * k)) & 7u];
            if (S.p == 0) {
                const KetDesc* __restrict__ kd = skets + S.g;
                if (kd->shift_kind == 0) {       // ZZ shift gate exp(i sigma alpha z0 z1); never both operands in J
                    const int t_id = S.t_id;
                    const size_t tbase = ((size_t)(t_id & ((1 << T.tid_lo_bits) - 1)) << T.a) | ((size_t)(t_id >> T.tid_lo_bits) << T.high_end);
                    const size_t xJ = tbase + ((size_t)(iJ & T.lowmask) | ((size_t)(iJ >> T.a) << T.start));
                    const int tb0 = tile_bit_of(kd->sb0), tb1 = tile_bit_of(kd->sb1);
                    const double sigma = kd->sigma;
                    const int j0b = type == 0 ? G0::jslot(tb0) : G1::jslot(tb0);
                    const int j1b = type == 0 ? G0::jslot(tb1) : G1::jslot(tb1);
                    const double z0 = ((xJ >> kd->sb0) & 1) ? -1.0 : 1.0;
                    const double z1 = ((xJ >> kd->sb1) & 1) ? -1.0 : 1.0;
                    // with the J operand at 0 (z = +1) the factor is exp(i sigma alpha z_other)
                    phi = cmul(phi, make_double2(A.ca, sigma * A.sa * z0 * z1));
                    const int jb = j0b >= 0 ? j0b : j1b;
                    if (jb >= 0) {               // flipping that J bit multiplies by exp(-2 i sigma alpha z_other)
                        const double zo = j0b >= 0 ? z1 : z0;
                        const c128 f = make_double2(A.c2a, -sigma * A.s2a * zo);
#pragma unroll
                        for (int k = 0; k < 5; ++k)
                            if (k == jb) F[k] = cmul(F[k], f);
                    }
                }
            }
#pragma unroll
            for (int b4 = 0; b4 < 2; ++b4) {
                const c128 p4 = b4 ? cmul(phi, F[4]) : phi;
#pragma unroll
                for (int b3 = 0; b3 < 2; ++b3) {
                    const c128 p3 = b3 ? cmul(p4, F[3]) : p4;
#pragma unroll
                    for (int b2 = 0; b2 < 2; ++b2) {
                        const c128 p2 = b2 ? cmul(p3, F[2]) : p3;
#pragma unroll
                        for (int b1 = 0; b1 < 2; ++b1) {
                            const c128 p1 = b1 ? cmul(p2, F[1]) : p2;
                            const int j = (b4 << 4) | (b3 << 3) | (b2 << 2) | (b1 << 1);
                            v[j] = cmul(v[j], p1);
                            v[j | 1] = cmul(v[j | 1], cmul(p1, F[0]));
                        }
                    }
                }
            }
            if (AJ) {
                DQ_FOR_REGS(v[j] = cmul(v[j], P.aj[j]))
            }
        } else {
            if (type == 0) {
                DQ_FOR_REGS(v[j] = tile[DQ_A_K0(j)])
            } else {
                DQ_FOR_REGS(v[j] = tile[DQ_A_K1(j)])
            }
        }
        // ---- stages 0..3 -------------------------------------------------------------------------------------------
        lift_bit<0>(v, P.rot[r][0]);
        lift_bit<1>(v, P.rot[r][1]);
        lift_bit<2>(v, P.rot[r][2]);
        lift_bit<3>(v, P.rot[r][3]);
        // ---- tail: stage 4 + the stores of the following exchange ----------------------------------------------------
        const double2 rc4 = P.rot[r][4];
        if (r == 0) {                                // K -> J exchange
            if (type == 0) {
                DQ_LIFT4_ST(DQ_A_K0)
                __syncwarp();                        // L: the exchange never leaves the warp's 1024 amplitudes
            } else {
                __syncwarp();                        // from here on the tile lives in the 128-byte pattern
                DQ_LIFT4_ST(DQ_A_K1)
                team_sync(team);
            }
        } else if (r == 1) {
            lift_bit<4>(v, rc4);
        } else if (r == 2) {                         // J -> K exchange
            if (type == 0) {
                DQ_LIFT4_ST(DQ_A_J0)
                __syncwarp();
            } else {
                DQ_LIFT4_ST(DQ_A_J1)
                team_sync(team);
            }
        } else {
            if (flags_all & F_STORE) {               // back into the landing layout of this buffer; one bulk store takes it from there
                if (type == 0) {
                    DQ_LIFT4_ST(DQ_A_K0)
                } else {
                    __syncwarp();                    // 64-byte pattern: these are not the slots this thread has just read
                    DQ_LIFT4_ST(DQ_A_LAND1)
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            } else {
                lift_bit<4>(v, rc4);
            }
        }
    }
#undef DQ_LIFT4_ST
#else
#pragma unroll 1
    for (int r = 0; r < 4; ++r) {
        // opaque per round: what a round derives from the thread index (slot addresses, gather indices, table addresses) is
        // formed where it is used instead of being hoisted out of the loop and kept alive across the butterfly run
        asm volatile("" : "+r"(tid));
        const int iK = type == 0 ? G0::baseK(tid) : G1::baseK(tid);
        const int iJ = type == 0 ? G0::baseJ(tid) : G1::baseJ(tid);
        const int sK = swz(iK), sJ = swz(iJ);
        if (r == 0) {
            // ---- KA : landing layout -> registers --------------------------------------------------------
            if (type == 0) {
                DQ_FOR_REGS(v[j] = tile[slot(sK, G0::regK(j))])
            } else {
                // an H tile whose rows are 64 bytes (n = 20) lands in the 64-byte swizzle (see process_tile)
                const int b0 = A.h_sw64 ? (iK ^ ((iK >> 3) & 3)) : sK;
                const int b1 = A.h_sw64 ? b0 : (b0 ^ 4);
                DQ_FOR_REGS(v[j] = tile[((j & 1) ? b1 : b0) + G1::regK(j)])
            }
        } else if (r == 1) {
            // ---- K -> J exchange --------------------------------------------------------------------------
            if (type == 0) {                     // L: the exchange never leaves the warp's 1024 amplitudes
                DQ_FOR_REGS(tile[slot(sK, G0::regK(j))] = v[j])
                __syncwarp();
                DQ_FOR_REGS(v[j] = tile[slot(sJ, G0::regJ(j))])
            } else {
                __syncwarp();                    // from here on the tile lives in the 128-byte pattern
                DQ_FOR_REGS(tile[slot(sK, G1::regK(j))] = v[j])
                team_sync(team);
                DQ_FOR_REGS(v[j] = tile[slot(sJ, G1::regJ(j))])
            }
        } else if (r == 2) {
            // ---- phase of the new step (the J1 rotations are done) ------------------------------------------
            const unsigned idx = ws.thread_idx[type][tid] | S.tile_idx;          // ten 3-bit table indices (ws_pack_idx)
            const int kb = type == 0 ? G0::kbits(iJ) : G1::kbits(iJ);
            c128 phi = cmul(S.tc[(iJ >> (type == 0 ? G0::spare_shift : G1::spare_shift)) & 3], P.tkk[kb]);
#pragma unroll
            for (int m = 0; m < 5; ++m) {
                if (T.xk_msk[m][0]) {            // launch-uniform
                    c128 w = P.xk[m][(idx >> (15 + 3 * m)) & 7u];
                    if ((kb >> m) & 1) w.y = -w.y;
                    phi = cmul(phi, w);
                }
            }
            c128 F[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) F[k] = P.fj[k][(idx >> (3 * k)) & 7u];
            if (S.p == 0) {
                const KetDesc* __restrict__ kd = skets + S.g;
                if (kd->shift_kind == 0) {       // ZZ shift gate exp(i sigma alpha z0 z1); never both operands in J
                    const int t_id = S.t_id;
                    const size_t tbase = ((size_t)(t_id & ((1 << T.tid_lo_bits) - 1)) << T.a) | ((size_t)(t_id >> T.tid_lo_bits) << T.high_end);
                    const size_t xJ = tbase + ((size_t)(iJ & T.lowmask) | ((size_t)(iJ >> T.a) << T.start));
                    const int tb0 = tile_bit_of(kd->sb0), tb1 = tile_bit_of(kd->sb1);
                    const double sigma = kd->sigma;
                    const int j0b = type == 0 ? G0::jslot(tb0) : G1::jslot(tb0);
                    const int j1b = type == 0 ? G0::jslot(tb1) : G1::jslot(tb1);
                    const double z0 = ((xJ >> kd->sb0) & 1) ? -1.0 : 1.0;
                    const double z1 = ((xJ >> kd->sb1) & 1) ? -1.0 : 1.0;
                    // with the J operand at 0 (z = +1) the factor is exp(i sigma alpha z_other)
                    phi = cmul(phi, make_double2(A.ca, sigma * A.sa * z0 * z1));
                    const int jb = j0b >= 0 ? j0b : j1b;
                    if (jb >= 0) {               // flipping that J bit multiplies by exp(-2 i sigma alpha z_other)
                        const double zo = j0b >= 0 ? z1 : z0;
                        const c128 f = make_double2(A.c2a, -sigma * A.s2a * zo);
#pragma unroll
                        for (int k = 0; k < 5; ++k)
                            if (k == jb) F[k] = cmul(F[k], f);
                    }
                }
            }
#pragma unroll
            for (int b4 = 0; b4 < 2; ++b4) {
                const c128 p4 = b4 ? cmul(phi, F[4]) : phi;
#pragma unroll
                for (int b3 = 0; b3 < 2; ++b3) {
                    const c128 p3 = b3 ? cmul(p4, F[3]) : p4;
#pragma unroll
                    for (int b2 = 0; b2 < 2; ++b2) {
                        const c128 p2 = b2 ? cmul(p3, F[2]) : p3;
#pragma unroll
                        for (int b1 = 0; b1 < 2; ++b1) {
                            const c128 p1 = b1 ? cmul(p2, F[1]) : p2;
                            const int j = (b4 << 4) | (b3 << 3) | (b2 << 2) | (b1 << 1);
                            v[j] = cmul(v[j], p1);
                            v[j | 1] = cmul(v[j | 1], cmul(p1, F[0]));
                        }
                    }
                }
            }
            if (AJ) {
                DQ_FOR_REGS(v[j] = cmul(v[j], P.aj[j]))
            }
        } else {
            // ---- J -> K exchange --------------------------------------------------------------------------
            if (type == 0) {                     // L: warp-local again (the warp owns t10, t11 in both rounds)
                DQ_FOR_REGS(tile[slot(sJ, G0::regJ(j))] = v[j])
                __syncwarp();
                DQ_FOR_REGS(v[j] = tile[slot(sK, G0::regK(j))])
            } else {
                DQ_FOR_REGS(tile[slot(sJ, G1::regJ(j))] = v[j])
                team_sync(team);
                DQ_FOR_REGS(v[j] = tile[slot(sK, G1::regK(j))])
            }
        }
#if DQ_MATH_LOCK
        if (r != 2) math_acquire(ws, tid >> 5);      // r == 2 still holds it from the J1 run (J1, phase, J2 are one math section)
#endif
        if (r == 1) WS_TRACE(A, S.item, 3);
        if (r == 3) WS_TRACE(A, S.item, 5);
        lift_run(v, P.rot[r]);
        if (r == 0) WS_TRACE(A, S.item, 2);
        if (r == 2) WS_TRACE(A, S.item, 4);
        if (r == 3) WS_TRACE(A, S.item, 6);
#if DQ_MATH_LOCK
        if (r != 1) math_release(ws, tid >> 5);
#endif
    }
#endif
    {
        const int flags = P.flags;
        const int iK = type == 0 ? G0::baseK(tid) : G1::baseK(tid);
#if DQ_WS_DIRECT_STORE
        if (flags & F_STORE) {
            // Straight from the registers to global memory: a lane's run of a warp-wide 16-byte store is 128 contiguous bytes
            // (L: t0..t2) or the tile's contiguous run (H: 64 bytes at n = 20), i.e. whole sectors either way.  The tile does
            // not go back through shared memory (one write + the bulk store's read of all 64 KB: a quarter of the kernel's
            // shared-memory wavefronts, the pipe that bounds it together with FP64), and the buffer is free for the loader
            // as soon as this team arrives.  L2 is asked to keep the lines (the next pass reads them back).
            const KetDesc* __restrict__ kd = skets + S.g;
            const int t_id = S.t_id;
            const size_t tbase = ((size_t)(t_id & ((1 << T.tid_lo_bits) - 1)) << T.a) | ((size_t)(t_id >> T.tid_lo_bits) << T.high_end);
            const size_t xK = tbase + ((size_t)(iK & T.lowmask) | ((size_t)(iK >> T.a) << T.start));
            c128* __restrict__ dst = kd->buf + xK;
            unsigned long long pol;
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
#pragma unroll
            for (int j = 0; j < kRegs; ++j)
                asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;"
                             ::"l"(dst + T.offK[j]), "d"(v[j].x), "d"(v[j].y), "l"(pol) : "memory");
            // publish order: these stores, this thread's arrival on `done` (release, cta), the storer's wait on it (acquire),
            // the storer's gpu-scope fence (cumulative over everything it has observed), the ket counter.  No fence here: it
            // would hold this warp until its 32 stores are acknowledged by L2 (measured: 76.7 vs 84.4 samples/s).
        }
#elif DQ_WS_PIPE
        // the tile went back into the landing layout with the last butterfly stage (round 3 tail above)
#else
        if (flags & F_STORE) {               // back into the landing layout of this buffer; one bulk store takes it from there
            const int sK = swz(iK);
            if (type == 0) {
                DQ_FOR_REGS(tile[slot(sK, G0::regK(j))] = v[j])
            } else {
                __syncwarp();                // 64-byte pattern: these are not the slots this thread has just read
                const int b0 = A.h_sw64 ? (iK ^ ((iK >> 3) & 3)) : sK;
                const int b1 = A.h_sw64 ? b0 : (b0 ^ 4);
                DQ_FOR_REGS(tile[((j & 1) ? b1 : b0) + G1::regK(j)] = v[j])
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
#endif
        if (flags & F_ENERGY) {
            const KetDesc* __restrict__ kd = skets + S.g;
            const int t_id = S.t_id;
            const size_t tbase = ((size_t)(t_id & ((1 << T.tid_lo_bits) - 1)) << T.a) | ((size_t)(t_id >> T.tid_lo_bits) << T.high_end);
            const size_t xK = tbase + ((size_t)(iK & T.lowmask) | ((size_t)(iK >> T.a) << T.start));
            const double* __restrict__ md = A.mdiag + xK;
            const c128* __restrict__ cr = (CROSS && kd->cross) ? kd->cross + xK : nullptr;
            double e = 0.0, e2 = 0.0;
#pragma unroll
            for (int j = 0; j < kRegs; ++j) {
                const size_t o = (size_t)T.offK[j];
                const double m = __ldg(md + o);
                e = fma(m, fma(v[j].x, v[j].x, v[j].y * v[j].y), e);
                if (CROSS && cr) {           // Re conj(a) ket
                    const c128 a0 = __ldcg(cr + o);
                    e2 = fma(m, fma(a0.x, v[j].x, a0.y * v[j].y), e2);
                }
            }
            for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
            if ((tid & 31) == 0) ws.red[b][tid >> 5] = e;
            if (CROSS) {
                for (int o = 16; o > 0; o >>= 1) e2 += __shfl_xor_sync(0xffffffffu, e2, o);
                if ((tid & 31) == 0) ws.red2[b][tid >> 5] = e2;
            }
        }
    }
}
#undef DQ_FOR_REGS

template <bool AJ, bool CROSS>
__global__ void __launch_bounds__(kWsThreads, 1) k_fused_ws(const __grid_constant__ LaunchArgs A) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);      // TMA swizzle: 1 KiB aligned buffers
    c128* tiles = reinterpret_cast<c128*>(smem_raw);
    PassStep* tables = reinterpret_cast<PassStep*>(smem_raw + sizeof(c128) * kTile * kBufs);     // one per tile buffer
    KetDesc* skets = reinterpret_cast<KetDesc*>(smem_raw + sizeof(c128) * kTile * kBufs + kBufs * sizeof(PassStep));
    __shared__ __align__(16) WsShared ws;

    if (threadIdx.x == 0) {
        for (int b = 0; b < kBufs; ++b) {
            mbar_init(&ws.full[b], 1);
            mbar_init(&ws.done[b], kTeamThreads);
            mbar_init(&ws.empty[b], 1);
        }
        for (int w = 0; w < kTeamThreads / 32; ++w) ws.math_lock[w] = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const int4* src = reinterpret_cast<const int4*>(A.kets);
        int4* dst = reinterpret_cast<int4*>(skets);
        for (int i = threadIdx.x; i < A.n_kets * (int)(sizeof(KetDesc) / 16); i += kWsThreads) dst[i] = __ldg(src + i);
    }
    if (threadIdx.x < kTeamThreads) {             // the same for both teams
        const int tid = threadIdx.x;
#pragma unroll
        for (int type = 0; type < 2; ++type) {
            const TypeGeom& T = A.geom[type];
            const int iJ = type == 0 ? Geo<0>::baseJ(tid) : Geo<1>::baseJ(tid);
            ws.thread_idx[type][tid] = ws_pack_idx(T, (size_t)(iJ & T.lowmask) | ((size_t)(iJ >> T.a) << T.start));
        }
    }
    __syncthreads();                              // the last CTA-wide barrier: the roles part ways here
    const unsigned total = ((unsigned)A.max_pass * (unsigned)A.group * (unsigned)A.n_groups) << A.ipp_log2;
    const int wg = threadIdx.x / kTeamThreads;

    if (wg < kTeams) {
        // ------------------------------ consumer team ---------------------------------------------------------
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kWsConsumerRegs));
        const int team = wg;
        for (unsigned k = (unsigned)team;; k += kTeams) {
            const int b = (int)(k % kBufs);
#if DQ_TRACE
            const long long t_top = clock64();
#endif
            mbar_wait(&ws.full[b], (k / kBufs) & 1u);
            if (ws.slot[b].stop) {
                mbar_arrive(&ws.done[b]);         // the storer sees the stop in tile order
                break;
            }
            PassStep& P = tables[b];
            c128* tile = tiles + (size_t)b * kTile;
#if DQ_TRACE
            if ((threadIdx.x & 31) == 0 && A.trace) {
                long long* tr = A.trace + (size_t)ws.slot[b].item * 48;
                tr[((threadIdx.x >> 5) & 3) * 8 + 0] = t_top;
                tr[((threadIdx.x >> 5) & 3) * 8 + 1] = clock64();
                if ((threadIdx.x & 127) == 0) { tr[32 + 2] = P.type; tr[32 + 6] = blockIdx.x * kTeams + team + 1; tr[32 + 7] = ws.slot[b].p; }
            }
#endif
            ws_tile<AJ, CROSS>(A, ws, skets, P, tile, b, team);
            mbar_arrive(&ws.done[b]);             // release: this thread's tile writes (fenced for the async proxy) and partials
            WS_TRACE(A, ws.slot[b].item, 7);
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kWsProducerRegs));
        const int w = (threadIdx.x - kTeams * kTeamThreads) >> 5;
        if (w == 0 && (threadIdx.x & 31) == 0) {
            // ------------------------------ loader ------------------------------------------------------------
            unsigned q_next = take_item(A.counters);
            unsigned stops = 0;
            for (unsigned k = 0;; ++k) {
                ItemInfo I;
                bool have = false;
                if (!stops) {
                    for (;;) {                    // next item that exists (ragged groups: kets with fewer passes are skipped)
                        const unsigned item = q_next;
                        if (item >= total) break;
                        q_next = take_item(A.counters);
                        decode_item(A, item, I);
                        if (item_valid(A, skets, I, total)) { have = true; break; }
                    }
                }
                const int b = (int)(k % kBufs);
                // first poll of the item's dependency BEFORE the buffer wait: its L2 round trip overlaps the time the buffer
                // is still busy instead of adding to the buffer's turn-around (relaxed, see ld_relaxed: what depends on the
                // value is the TMA load below, which reads L2)
                int ctr = 0;
                unsigned need = 0, seen = 0;
                const bool dep = have && item_dependency(A, skets, I.p, I.g, ctr, need);
                if (dep) seen = ld_relaxed(&A.counters[ctr]);
                if (k >= kBufs) mbar_wait(&ws.empty[b], ((k / kBufs) - 1u) & 1u);   // the store of tile k - 3 has read the buffer
                WsSlot& S = ws.slot[b];
                if (!have) {                      // out of work: one stop per team, then done
                    S.stop = 1;
                    mbar_arrive(&ws.full[b]);
                    if (++stops == kTeams) break;
                    continue;
                }
                if (dep && seen < need)
                    while (ld_relaxed(&A.counters[ctr]) < need) __nanosleep(20);
                const KetDesc* __restrict__ kd = skets + I.g;
                const PassStep* ps = kd->steps + I.p;
                S.p = I.p;
                S.g = I.g;
                S.t_id = I.t_id;
                S.grp = I.grp;
                S.item = I.item;
                S.stop = 0;
                const int type = (I.p + kd->cls) & 1;
                {
                    const TypeGeom& T = A.geom[type];
                    S.tile_idx = ws_pack_idx(T, ((size_t)(I.t_id & ((1 << T.tid_lo_bits) - 1)) << T.a) | ((size_t)(I.t_id >> T.tid_lo_bits) << T.high_end));
                }
                const CUtensorMap* map = A.maps + 2 * (I.p == 0 ? kd->map_src : kd->map_buf) + type;
                c128* tile = tiles + (size_t)b * kTile;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&ws.full[b], (unsigned)(sizeof(c128) * kTile + sizeof(PassStep) + sizeof(S.tc)));
                if (type == 0) {
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                                 ::"r"(smem_u32(tile)), "l"(map), "r"(smem_u32(&ws.full[b])), "r"(0), "r"(0), "r"(2 * I.t_id) : "memory");
                } else {
                    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                                 ::"r"(smem_u32(tile)), "l"(map), "r"(smem_u32(&ws.full[b])), "r"(0), "r"(I.t_id << A.h_c1_shift), "r"(0), "r"(0) : "memory");
                }
                bulk_copy_g2s(&tables[b], ps, (unsigned)sizeof(PassStep), &ws.full[b]);
                const double2* tc = A.tc + ((unsigned long long)(ps - A.steps0) << A.n_col_bits) + ((unsigned)I.t_id << 2);
                bulk_copy_g2s(S.tc, tc, (unsigned)sizeof(S.tc), &ws.full[b]);
            }
        } else if (w == 1 && (threadIdx.x & 31) == 0) {
            // ------------------------------ storer ------------------------------------------------------------
            int pend_g = -1;                      // tile whose completion is not published yet
            bool pend_stored = false;
            unsigned stops = 0;
            for (unsigned k = 0;; ++k) {
                const int b = (int)(k % kBufs);
                const unsigned par = (k / kBufs) & 1u;
                while (!mbar_test(&ws.done[b], par)) {
                    if (pend_g >= 0) {            // idle: publish what is pending (its store is the most recent group)
                        if (pend_stored) bulk_wait_group<0>();
                        atomicAdd(&A.counters[1 + pend_g], 1u);
                        pend_g = -1;
                    } else {
                        __nanosleep(20);
                    }
                }
                const WsSlot& S = ws.slot[b];
                if (S.stop) {
                    if (++stops == kTeams) break;
                    continue;
                }
                const KetDesc* __restrict__ kd = skets + S.g;
                const int flags = tables[b].flags;
                const int type = tables[b].type;
#if DQ_WS_DIRECT_STORE
                (void)type;
                const bool stored = false;        // the consumers stored (and fenced) the tile themselves: only the publishing is left
#else
                const bool stored = (flags & F_STORE) != 0;
#endif
                if (stored) {
                    store_tile(A, kd, type, S.t_id, tiles + (size_t)b * kTile);
                    // the tile before this one: its store is now the second most recent group
                    if (pend_g >= 0) {
                        if (pend_stored) bulk_wait_group<1>();
                        atomicAdd(&A.counters[1 + pend_g], 1u);
                        pend_g = -1;
                    }
                    bulk_wait_read();             // the store has read the buffer: the loader may refill it
                } else if (pend_g >= 0) {
                    if (pend_stored) bulk_wait_group<0>();
                    atomicAdd(&A.counters[1 + pend_g], 1u);
                    pend_g = -1;
                }
                if (flags & F_ENERGY) {
                    if (kd->partial) kd->partial[S.grp] = (ws.red[b][0] + ws.red[b][1] + ws.red[b][2] + ws.red[b][3]) * kd->escale;
                    if (CROSS && kd->cross) kd->partial2[S.grp] = (ws.red2[b][0] + ws.red2[b][1] + ws.red2[b][2] + ws.red2[b][3]) * kd->escale2;
                }
#if DQ_WS_DIRECT_STORE
                const int g_done = S.g;           // the slot is the loader's again once `empty` is signalled
                mbar_arrive(&ws.empty[b]);        // nothing reads the buffer any more: the loader may refill it right away
                // the consumers' stores happen-before their arrival on `done`; this fence makes them visible at gpu scope
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                atomicAdd(&A.counters[1 + g_done], 1u);
#else
                pend_g = S.g;
                pend_stored = stored;
                mbar_arrive(&ws.empty[b]);
#endif
            }
            if (pend_g >= 0) {
                bulk_wait_group<0>();
                atomicAdd(&A.counters[1 + pend_g], 1u);
            }
        }
    }
}

#ifdef DQ_PROBE_ONLY
// Register-allocation probe (build with -DDQ_PROBE_ONLY, never linked): the consumer loop alone under the register budget
// the consumer warpgroups get from setmaxnreg (launch bounds 280 threads -> 232 registers).
template <bool AJ, bool CROSS>
__global__ void k_ws_probe(const __grid_constant__ LaunchArgs A) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    c128* tiles = reinterpret_cast<c128*>(smem_raw);
    PassStep* tables = reinterpret_cast<PassStep*>(smem_raw + sizeof(c128) * kTile * kBufs);
    KetDesc* skets = reinterpret_cast<KetDesc*>(smem_raw + sizeof(c128) * kTile * kBufs + kBufs * sizeof(PassStep));
    __shared__ __align__(16) WsShared ws;
    const int team = threadIdx.x / kTeamThreads;
    for (unsigned k = (unsigned)team;; k += kTeams) {
        const int b = (int)(k % kBufs);
        mbar_wait(&ws.full[b], (k / kBufs) & 1u);
        if (ws.slot[b].stop) {
            mbar_arrive(&ws.done[b]);
            break;
        }
        PassStep& P = tables[b];
        c128* tile = tiles + (size_t)b * kTile;
        ws_tile<AJ, CROSS>(A, ws, skets, P, tile, b, team);
        mbar_arrive(&ws.done[b]);
    }
}
template __global__ void k_ws_probe<false, false>(const __grid_constant__ LaunchArgs);
}  // namespace fused
}  // namespace dq
#else
// ------------------------------------------------------------------------------------------
// table setup: one CTA per pass-step (y = column-table chunk)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double zsign(int bits, int i) { return ((bits >> i) & 1) ? -1.0 : 1.0; }

__global__ void __launch_bounds__(128) k_setup(const SetupJob* __restrict__ jobs, const double* __restrict__ rows,
                                               int row_len, int n_zz, const TypePlan* __restrict__ types,
                                               PassStep* __restrict__ steps, double2* __restrict__ tc,
                                               int n_col_bits) {
    const SetupJob job = jobs[blockIdx.x];
    const bool scaled = (job.flags & F_SCALED) != 0;
    const TypePlan& T = types[job.type];
    PassStep& P = steps[blockIdx.x];
    const double* pre = job.row_pre >= 0 ? rows + job.row_pre * row_len : nullptr;
    const double* cur = job.row_cur >= 0 ? rows + job.row_cur * row_len : nullptr;
    const int off_x = 1 + n_zz;
    const unsigned long long tc_off = (unsigned long long)blockIdx.x << n_col_bits;
    const int tid = threadIdx.x;

    // rotation angles by slot {KA, J1, J2, KB}; the final pass (no `cur`) moves the K rotations of the
    // last step to slot KB so that the energy is reduced from the coalesced outer layout
    double ang[4][5];
    double scale = 1.0;
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const double ka = (pre && T.kq[b] >= 0) ? pre[off_x + T.kq[b]] : 0.0;
        const double j1 = (pre && T.jq[b] >= 0) ? pre[off_x + T.jq[b]] : 0.0;
        const double j2 = (cur && T.jq[b] >= 0) ? cur[off_x + T.jq[b]] : 0.0;
        const double kbv = (cur && T.kq[b] >= 0) ? cur[off_x + T.kq[b]] : 0.0;
        // final pass (no `cur`): the straight-line body moves the K rotations of the last step to slot KB (energy from the
        // coalesced outer layout either way); the lifting body keeps them in KA so that every pending factor is resolved
        // by this pass's phase and its KB round is the identity
        ang[0][b] = (cur || scaled) ? ka : 0.0;
        ang[1][b] = j1;
        ang[2][b] = j2;
        ang[3][b] = cur ? kbv : (scaled ? 0.0 : ka);
    }
    // Lifting butterflies (scaled mode) leave amplitude x with the factor  prod_b (bit_b(x) ? cos_b : 1/cos_b)  over every
    // bit b rotated since the last phase: the bits of THIS pass type by its KA / J1 rounds and the bits of the OTHER type by
    // the previous pass's J2 / KB rounds -- together every qubit, each with its angle of row `pre`.  The phase of this pass
    // "resolves" them: the inverse factor is a product over index bits, so it folds into the tables by position of the bit
    // (K position -> tkk, J position -> base and fj, column bit -> tc).
    double rk[5], rj[5];
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        rk[b] = (scaled && pre && T.kq_all[b] >= 0) ? cos(pre[off_x + T.kq_all[b]]) : 1.0;
        rj[b] = (scaled && pre && T.jq_all[b] >= 0) ? cos(pre[off_x + T.jq_all[b]]) : 1.0;
    }
    (void)scale;
    if (blockIdx.y == 0) {
        if (tid < 20) {
            const int s = tid / 5, b = tid % 5;
            double sn, cs;
            sincos(ang[s][b], &sn, &cs);
            P.rot[s][b] = scaled ? make_double2(sn / cs, sn * cs) : make_double2(cs, sn);     // (tan, sin cos) | (cos, sin)
        }
        if (tid == 0) {
            P.flags = job.flags & ~F_SCALED;
            P.type = job.type;
            P.tc_offset = tc_off;
        }
        // phase tables; the final pass has no phase: every angle is 0 and only the cosine scale remains
        if (tid < 32) {                         // tkk[kb] and aj[j]
            double a = cur ? cur[0] : 0.0, aa = 0.0;
            if (cur)
                for (int e = 0; e < T.n_pairs; ++e) {
                    const double g = cur[1 + e];
                    switch (T.cls[e]) {
                        case 0: a += g; aa += g * (zsign(tid, T.i0[e]) * zsign(tid, T.i1[e]) - 1.0); break;
                        case 1: a += g * zsign(tid, T.i0[e]) * zsign(tid, T.i1[e]); break;
                        case 2: a += g * zsign(tid, T.i1[e]); break;
                        default: break;
                    }
                }
            double res = 1.0;                   // resolve: K positions by their bit of tid, J positions at bit 0
#pragma unroll
            for (int b = 0; b < 5; ++b) res *= (((tid >> b) & 1) ? 1.0 / rk[b] : rk[b]) * rj[b];
            double sn, cs;
            sincos(a, &sn, &cs);
            P.tkk[tid] = make_double2(res * cs, -res * sn);
            sincos(aa, &sn, &cs);
            P.aj[tid] = make_double2(cs, -sn);
        } else if (tid < 32 + 40) {             // xk[m][pat]
            const int m = (tid - 32) >> 3, pat = (tid - 32) & 7;
            double a = 0.0;
            if (cur)
                for (int t = 0; t < kMaxNbr; ++t)
                    if (T.g.xk_msk[m][t]) a += cur[1 + T.xk_pair[m][t]] * zsign(pat, t);
            double sn, cs;
            sincos(a, &sn, &cs);
            P.xk[m][pat] = make_double2(cs, -sn);
        } else if (tid < 32 + 80) {             // fj[k][pat] = exp(+2i w_k)
            const int k = (tid - 72) >> 3, pat = (tid - 72) & 7;
            double a = 0.0;
            if (cur)
                for (int t = 0; t < kMaxNbr; ++t)
                    if (T.g.fj_msk[k][t]) a += cur[1 + T.fj_pair[k][t]] * zsign(pat, t);
            double sn, cs;
            sincos(2.0 * a, &sn, &cs);
            const double res = 1.0 / (rj[k] * rj[k]);           // resolve: J position k at bit 1 relative to bit 0
            P.fj[k][pat] = make_double2(res * cs, res * sn);
        }
    }
    // column table: CC pairs + fields from J neighbours
    const unsigned ncol = 1u << n_col_bits;
    for (unsigned col = blockIdx.y * blockDim.x + tid; col < ncol; col += gridDim.y * blockDim.x) {
        double a = 0.0;
        if (cur)
            for (int e = 0; e < T.n_pairs; ++e) {
                const double g = cur[1 + e];
                if (T.cls[e] == 5) a += g * ((((col >> T.i0[e]) ^ (col >> T.i1[e])) & 1) ? -1.0 : 1.0);
                else if (T.cls[e] == 3) a += g * (((col >> T.i1[e]) & 1) ? -1.0 : 1.0);
            }
        double res = 1.0;                       // resolve: column bits
        if (scaled && pre)
            for (int i = 0; i < n_col_bits; ++i)
                if (T.colq[i] >= 0) {
                    const double c = cos(pre[off_x + T.colq[i]]);
                    res *= ((col >> i) & 1) ? 1.0 / c : c;
                }
        double sn, cs;
        sincos(a, &sn, &cs);
        tc[tc_off + col] = make_double2(res * cs, -res * sn);
    }
}

__global__ void k_sum_partials(const double* __restrict__ partial, int tiles, int count, const int* __restrict__ out_index,
                               double* __restrict__ out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < count; i += 32) acc += partial[(size_t)blockIdx.x * tiles + i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) out[out_index[blockIdx.x]] = acc;
}

// linear mode: out[b][2i] holds E+ and out[b][2i+1] the cross term C = Re <a|M|ket+>; ket- = 2a/sqrt(1+r^2) - ket+
// gives E- = 4 Ea/(1+r^2) + E+ - 4 C/sqrt(1+r^2)   (sim_plain.py:197-220 by linearity of the suffix evolution)
__global__ void k_linear_fix(double* __restrict__ out, const double* __restrict__ ea, int kets_per, int total, double r) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total || (i & 1) == 0) return;
    const int b = i / kets_per;
    const double inv = 1.0 / (1.0 + r * r);
    out[i] = 4.0 * ea[b] * inv + out[i - 1] - 4.0 * out[i] * sqrt(inv);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct Plan {
    bool ok = false;
    int n = 0, n_col_bits = 0, tiles_log2 = 0, sub_log2 = 0;
    bool has_aj = false;
    TypePlan types[2];
    int jphys[2][5];
    DevBuf d_types, jobs, steps, tc, kets, counters, partials, out_index, work, rows, phi, uniform, trace, afin, ea, ea_index;
    long long trace_items = 0;
    size_t smem_bytes = 0, ws_smem_bytes = 0;
    int ctas_per_sm = 0;
    int counter_slots = 0, counter_cursor = 0;
    std::vector<cudaEvent_t> ev;    // option time_launches: event pairs around every pass-kernel launch of the last run
    int ev_used = 0;
    // TMA views, one (L, H) pair per state buffer ever named by a ket; device copy refreshed before a launch
    std::vector<const void*> map_keys;
    std::vector<CUtensorMap> h_maps;
    DevBuf d_maps;
    size_t maps_uploaded = 0;
    int h_c1_shift = 0;
};
constexpr size_t kMaxMaps = 4096;   // buffer pairs cached per problem

static std::vector<std::pair<dq_ising*, Plan*>> g_plans;     // one plan per problem

static Plan* find_plan(const dq_ising* p) {
    for (auto& kv : g_plans)
        if (kv.first == p) return kv.second;
    return nullptr;
}

static bool build_type(const dq_ising* p, int type, TypePlan& T, int* jphys_out) {
    const int n = p->n;
    memset(&T, 0, sizeof(T));
    int a, b;
    if (type == 0) {            // L: tile = physical bits [0,12); K, J as Geo<0>
        a = kTileBits; T.start = kTileBits; b = 0;
        T.spare_shift = 10;
    } else {                    // H: tile = low spectators [0,a) + physical bits [10,n); K, J as Geo<1>
        b = n - 10;
        a = kTileBits - b; T.start = 10;
        T.spare_shift = 0;
    }
    const int* jbits = kJBits[type];
    const int* kbits = kKBits[type];
    auto spread = [](int j, const int* bits) { int t = 0; for (int i = 0; i < 5; ++i) t |= ((j >> i) & 1) << bits[i]; return t; };
    T.g.a = a;
    T.g.start = T.start;
    T.g.lowmask = (1 << a) - 1;
    T.g.tid_lo_bits = T.start - a;
    T.g.high_end = T.start + b;
    T.n_col_bits = n - 10;
    auto phys_of_tile_bit = [&](int t) { return t < a ? t : T.start + (t - a); };
    auto phys_off = [&](int i) { return (i & T.g.lowmask) | ((i >> a) << T.start); };
    auto tile_bit_of_phys = [&](int pos) { return pos < a ? pos : (pos >= T.start && pos < T.start + b ? pos - T.start + a : -1); };
    for (int j = 0; j < kRegs; ++j) {
        T.g.offK[j] = phys_off(spread(j, kbits));
        T.g.offJ[j] = phys_off(spread(j, jbits));
    }
    // which physical bits are rotated by this pass type
    auto active = [&](int pos) { return type == 0 ? (pos < 10) : (pos >= 10); };
    int qubit_of_pos[64];
    for (int q = 0; q < n; ++q) qubit_of_pos[p->bitpos[q]] = q;
    int jphys[5], kphys[5];
    for (int i = 0; i < 5; ++i) {
        jphys[i] = phys_of_tile_bit(jbits[i]);
        kphys[i] = phys_of_tile_bit(kbits[i]);
        jphys_out[i] = jphys[i];
        T.jq[i] = active(jphys[i]) ? qubit_of_pos[jphys[i]] : -1;
        T.kq[i] = active(kphys[i]) ? qubit_of_pos[kphys[i]] : -1;
        T.jq_all[i] = qubit_of_pos[jphys[i]];
        T.kq_all[i] = qubit_of_pos[kphys[i]];
    }
    // column bits: the two spare tile bits, then the tile-id bits in ascending physical order
    std::vector<int> colphys;
    colphys.push_back(phys_of_tile_bit(T.spare_shift));
    colphys.push_back(phys_of_tile_bit(T.spare_shift + 1));
    for (int pos = 0; pos < n; ++pos)
        if (tile_bit_of_phys(pos) < 0) colphys.push_back(pos);
    if ((int)colphys.size() != n - 10) return false;
    for (int i = 0; i < 16; ++i) T.colq[i] = i < (int)colphys.size() ? qubit_of_pos[colphys[i]] : -1;
    auto local = [&](int pos, int& kind, int& idx) {
        for (int i = 0; i < 5; ++i) if (jphys[i] == pos) { kind = 0; idx = i; return; }
        for (int i = 0; i < 5; ++i) if (kphys[i] == pos) { kind = 1; idx = i; return; }
        for (size_t i = 0; i < colphys.size(); ++i) if (colphys[i] == pos) { kind = 2; idx = (int)i; return; }
        kind = -1; idx = -1;
    };
    if (p->n_zz > kMaxPairs) return false;
    T.n_pairs = p->n_zz;
    int fj_cnt[5] = {0, 0, 0, 0, 0}, xk_cnt[5] = {0, 0, 0, 0, 0};
    for (int e = 0; e < p->n_zz; ++e) {
        int ka, ia, kb2, ib;
        local(p->pa[e], ka, ia);
        local(p->pb[e], kb2, ib);
        if (ka < 0 || kb2 < 0) return false;
        if (ka > kb2) { std::swap(ka, kb2); std::swap(ia, ib); }
        const int posb = (kb2 == 1) ? kphys[ib] : (kb2 == 2 ? colphys[ib] : jphys[ib]);
        if (ka == 0 && kb2 == 0) { T.cls[e] = 0; T.has_aj = 1; }
        else if (ka == 1 && kb2 == 1) T.cls[e] = 1;
        else if (ka == 0 && kb2 == 1) T.cls[e] = 2;
        else if (ka == 0 && kb2 == 2) T.cls[e] = 3;
        else if (ka == 1 && kb2 == 2) T.cls[e] = 4;
        else T.cls[e] = 5;
        T.i0[e] = (signed char)ia;
        T.i1[e] = (signed char)ib;
        if (ka == 0 && kb2 != 0) {              // neighbour of J bit ia outside J
            if (fj_cnt[ia] >= kMaxNbr) return false;
            const int t = fj_cnt[ia]++;
            T.g.fj_pos[ia][t] = posb; T.g.fj_msk[ia][t] = 1; T.fj_pair[ia][t] = e;
        }
        if (ka == 1 && kb2 == 2) {              // column neighbour of K bit ia
            if (xk_cnt[ia] >= kMaxNbr) return false;
            const int t = xk_cnt[ia]++;
            T.g.xk_pos[ia][t] = posb; T.g.xk_msk[ia][t] = 1; T.xk_pair[ia][t] = e;
        }
    }
    return true;
}

// ---- TMA views of a state buffer ----------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
        cudaGetLastError();
    }
    return fn;
}

// out[0]: L view, a tile = 4096 consecutive amplitudes as {8 amplitudes (128 B), 256 rows, 2 blocks};
// out[1]: H view, a tile = 2^a consecutive low amplitudes x every high index, as
//         {2^w amplitudes, 2^(a-w) of the 2^(10-w) low-column groups, high bits 10..17, high bits 18..}, w = min(a, 3).
// Both boxes are written to shared memory densely in this order = the natural tile order, 128-byte swizzle
// (64-byte swizzle when the H rows are only 64 bytes wide: a narrower row would be padded to the swizzle span).
static bool make_maps(int n, const void* base, CUtensorMap* out) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    const cuuint32_t ones[4] = {1, 1, 1, 1};
    {
        const cuuint64_t dims[3] = {16, 256, (cuuint64_t)1 << (n - 11)};
        const cuuint64_t strides[2] = {128, 128 * 256};
        const cuuint32_t box[3] = {16, 256, 2};
        if (enc(&out[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), dims, strides, box, ones,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    {
        const int a = 22 - n, w = std::min(a, 3);
        const int hi_lo = std::min(8, n - 10), hi_hi = std::max(0, n - 18);
        const cuuint64_t dims[4] = {(cuuint64_t)2 << w, (cuuint64_t)1 << (10 - w), (cuuint64_t)1 << hi_lo, (cuuint64_t)1 << hi_hi};
        const cuuint64_t strides[3] = {(cuuint64_t)16 << w, (cuuint64_t)16 << 10, (cuuint64_t)16 << 18};
        const cuuint32_t box[4] = {(cuuint32_t)2 << w, (cuuint32_t)1 << (a - w), (cuuint32_t)1 << hi_lo, (cuuint32_t)1 << hi_hi};
        if (enc(&out[1], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void*>(base), dims, strides, box, ones,
                CU_TENSOR_MAP_INTERLEAVE_NONE, w < 3 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    return true;
}

// Index of the view pair of `base` (created on first use); -1 when the driver refuses the encoding.
static int map_index(Plan* pl, const void* base) {
    for (size_t i = 0; i < pl->map_keys.size(); ++i)
        if (pl->map_keys[i] == base) return (int)i;
    CUtensorMap m[2];
    if (!make_maps(pl->n, base, m)) return -1;
    pl->map_keys.push_back(base);
    pl->h_maps.push_back(m[0]);
    pl->h_maps.push_back(m[1]);
    return (int)pl->map_keys.size() - 1;
}

// Make room for `need` more buffers; a full cache is dropped (indices are only meaningful within one run).
static int maps_begin(dq_ising* p, Plan* pl, size_t need) {
    DQ_REQUIRE(need <= kMaxMaps, "fused engine: more than %zu distinct state buffers in one call", kMaxMaps);
    if (pl->map_keys.size() + need > kMaxMaps) {
        DQ_CUDA(cudaStreamSynchronize(p->ctx->stream));
        pl->map_keys.clear();
        pl->h_maps.clear();
        pl->maps_uploaded = 0;
    }
    return DQ_OK;
}

static int maps_upload(dq_ising* p, Plan* pl) {
    if (pl->maps_uploaded == pl->h_maps.size()) return DQ_OK;
    DQ_CUDA(cudaMemcpyAsync(pl->d_maps.as<CUtensorMap>() + pl->maps_uploaded, pl->h_maps.data() + pl->maps_uploaded,
                            (pl->h_maps.size() - pl->maps_uploaded) * sizeof(CUtensorMap), cudaMemcpyHostToDevice, p->ctx->stream));
    pl->maps_uploaded = pl->h_maps.size();
    return DQ_OK;
}

template <typename K> static bool prep_kernel(K kern, size_t smem, int* occ, int threads = kThreads) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, threads, smem) == cudaSuccess && *occ >= 1;
}

static Plan* get_plan(dq_ising* p) {
    Plan* pl = find_plan(p);
    if (pl) return pl;
    pl = new Plan();
    g_plans.push_back({p, pl});
    pl->n = p->n;
    if (p->ctx->set_device() != DQ_OK) return pl;
    if (p->n < kTileBits || p->n > 20) return pl;       // two sets cover n <= 20
    for (int t = 0; t < 2; ++t)
        if (!build_type(p, t, pl->types[t], pl->jphys[t])) return pl;
    pl->has_aj = pl->types[0].has_aj || pl->types[1].has_aj;
    pl->n_col_bits = p->n - 10;
    pl->tiles_log2 = p->n - kTileBits;
    pl->smem_bytes = 1024 + sizeof(c128) * kTile * kBufs + 2 * kTeams * sizeof(PassStep) + kMaxGroup * sizeof(KetDesc);
    pl->ws_smem_bytes = 1024 + sizeof(c128) * kTile * kBufs + kBufs * sizeof(PassStep) + kMaxGroup * sizeof(KetDesc);
    {
        const int a = 22 - p->n;
        pl->h_c1_shift = a - std::min(a, 3);
    }
    if (!encode_fn()) return pl;
    if (pl->d_maps.reserve(kMaxMaps * 2 * sizeof(CUtensorMap)) != DQ_OK) return pl;
    if (pl->d_types.reserve(sizeof(TypePlan) * 2) != DQ_OK) return pl;
    if (cudaMemcpy(pl->d_types.p, pl->types, sizeof(TypePlan) * 2, cudaMemcpyHostToDevice) != cudaSuccess) return pl;
    int occ = 0, o2 = 0;
    int o3 = 0, o4 = 0;
    bool good = pl->has_aj ? (prep_kernel(k_fused_ws<true, false>, pl->ws_smem_bytes, &occ, kWsThreads) &&
                              prep_kernel(k_fused_passes<false, true, false>, pl->smem_bytes, &o2) &&
                              prep_kernel(k_fused_ws<true, true>, pl->ws_smem_bytes, &o3, kWsThreads) &&
                              prep_kernel(k_fused_passes<false, true, true>, pl->smem_bytes, &o4))
                           : (prep_kernel(k_fused_ws<false, false>, pl->ws_smem_bytes, &occ, kWsThreads) &&
                              prep_kernel(k_fused_passes<false, false, false>, pl->smem_bytes, &o2) &&
                              prep_kernel(k_fused_ws<false, true>, pl->ws_smem_bytes, &o3, kWsThreads) &&
                              prep_kernel(k_fused_passes<false, false, true>, pl->smem_bytes, &o4));
    o2 = std::min(o2, std::min(o3, o4));
    if (!good) { cudaGetLastError(); return pl; }
    pl->ctas_per_sm = std::min(occ, o2);
    pl->counter_slots = 1 << 16;
    if (pl->counters.reserve(pl->counter_slots * sizeof(unsigned)) != DQ_OK) return pl;
    pl->ok = true;
    return pl;
}

// A trajectory = rows [row0, row0 + n_steps); class c = index of its first pass type.
struct Traj {
    long long row0;
    int n_steps;
    int cls;
    bool final_energy;
    size_t step0;          // first PassStep index (filled by add_traj)
    bool final_both = false;   // final pass stores AND reduces (the unshifted trajectory of the linear mode)
    bool scaled = false;       // lifting-form tables (every |x angle| of the trajectory's sample <= 1)
};

static void add_traj(std::vector<SetupJob>& jobs, Traj& t) {
    t.step0 = jobs.size();
    for (int p = 0; p <= t.n_steps; ++p) {
        SetupJob j;
        j.row_pre = p >= 1 ? t.row0 + p - 1 : -1;
        j.row_cur = p < t.n_steps ? t.row0 + p : -1;
        j.type = (p + t.cls) & 1;
        j.flags = p < t.n_steps ? F_STORE : (t.final_both ? (F_ENERGY | F_STORE) : (t.final_energy ? F_ENERGY : F_STORE));
        if (t.scaled) j.flags |= F_SCALED;
        jobs.push_back(j);
    }
}

static int run_setup(dq_ising* p, Plan* pl, const std::vector<SetupJob>& jobs, const double* d_rows) {
    cudaStream_t st = p->ctx->stream;
    DQ_TRY(pl->jobs.reserve(jobs.size() * sizeof(SetupJob)));
    DQ_TRY(pl->steps.reserve(jobs.size() * sizeof(PassStep)));
    DQ_TRY(pl->tc.reserve((jobs.size() << pl->n_col_bits) * sizeof(double2)));
    DQ_CUDA(cudaMemcpyAsync(pl->jobs.p, jobs.data(), jobs.size() * sizeof(SetupJob), cudaMemcpyHostToDevice, st));
    const unsigned ncol = 1u << pl->n_col_bits;
    dim3 grid((unsigned)jobs.size(), std::max(1u, std::min(8u, ncol / 128)));
    k_setup<<<grid, 128, 0, st>>>(pl->jobs.as<SetupJob>(), d_rows, p->row_len, p->n_zz, pl->d_types.as<TypePlan>(),
                                  pl->steps.as<PassStep>(), pl->tc.as<double2>(), pl->n_col_bits);
    p->ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

#if DQ_TRACE
static long long all_items_for_trace(int n_kets, int tiles_log2, int max_pass) { return ((long long)n_kets << tiles_log2) * max_pass; }
#endif

static int launch_group(dq_ising* p, Plan* pl, const KetDesc* d_kets, int n_kets, int max_pass, bool scaled, double r,
                        int chain_group = 0) {
    cudaStream_t st = p->ctx->stream;
    DQ_REQUIRE(n_kets <= kMaxGroup, "fused engine: at most %d kets per launch", kMaxGroup);
    const int group = (chain_group > 0 && chain_group < n_kets) ? chain_group : n_kets;
    const int n_groups = (n_kets + group - 1) / group;
    if (pl->counter_cursor + 1 + n_kets > pl->counter_slots) pl->counter_cursor = 0;
    unsigned* ctr = pl->counters.as<unsigned>() + pl->counter_cursor;
    pl->counter_cursor += 1 + n_kets;
    DQ_CUDA(cudaMemsetAsync(ctr, 0, (1 + n_kets) * sizeof(unsigned), st));
    LaunchArgs A;
    A.trace = nullptr;
#if DQ_TRACE
    {
        const size_t need = (size_t)all_items_for_trace(group * n_groups, pl->tiles_log2, max_pass) * 48 * sizeof(long long);
        if (pl->trace.reserve(need) == DQ_OK) { cudaMemsetAsync(pl->trace.p, 0, need, st); A.trace = pl->trace.as<long long>(); pl->trace_items = all_items_for_trace(group * n_groups, pl->tiles_log2, max_pass); }
    }
#endif
    A.kets = d_kets;
    A.maps = pl->d_maps.as<CUtensorMap>();
    A.h_c1_shift = pl->h_c1_shift;
    A.h_sw64 = (22 - pl->n) < 3 ? 1 : 0;
    A.tc = pl->tc.as<double2>();
    A.steps0 = pl->steps.as<PassStep>();
    A.n_col_bits = pl->n_col_bits;
    A.mdiag = p->mdiag.as<double>();
    A.counters = ctr;
    A.n_kets = n_kets;
    A.group = group;
    A.n_groups = n_groups;
    A.max_pass = max_pass;
    A.inv_group = group > 1 ? ~0ull / (unsigned long long)group + 1ull : 0ull;
    A.inv_pass = max_pass > 1 ? ~0ull / (unsigned long long)max_pass + 1ull : 0ull;
    A.tiles_log2 = pl->tiles_log2;
    // multi-tile work items exist in the straight-line kernel only (measured slower, kept as an experiment switch)
    pl->sub_log2 = scaled ? 0 : std::min(std::max(0, p->item_tiles_log2), pl->tiles_log2);
    A.sub_log2 = pl->sub_log2;
    A.ipp_log2 = pl->tiles_log2 - pl->sub_log2;
    const double alpha = atan(r);
    A.r = r; A.ca = cos(alpha); A.sa = sin(alpha); A.c2a = cos(2 * alpha); A.s2a = sin(2 * alpha);
    A.rtau = r / (1.0 + r * r);
    A.geom[0] = pl->types[0].g;
    A.geom[1] = pl->types[1].g;
    const long long all_items = ((long long)(group * n_groups) << A.ipp_log2) * max_pass;
    long long grid = (long long)p->ctx->prop.multiProcessorCount * (p->grid_per_sm > 0 ? std::min(p->grid_per_sm, pl->ctas_per_sm) : pl->ctas_per_sm);
    if (grid > all_items) grid = all_items;
    const bool timed = p->time_launches != 0;
    if (timed) {
        while ((int)pl->ev.size() < pl->ev_used + 2) {
            cudaEvent_t e;
            DQ_CUDA(cudaEventCreate(&e));
            pl->ev.push_back(e);
        }
        DQ_CUDA(cudaEventRecord(pl->ev[pl->ev_used], st));
    }
#define DQ_LAUNCH(J, C) k_fused_passes<false, J, C><<<(unsigned)grid, kThreads, pl->smem_bytes, st>>>(A)
#define DQ_LAUNCH_LIFT(J, C) k_fused_ws<J, C><<<(unsigned)grid, kWsThreads, pl->ws_smem_bytes, st>>>(A)
    const bool cross = p->linear != 0;
    if (scaled) {                           // |angle| <= 1 everywhere: loop-form kernel (lifting butterflies)
        if (pl->has_aj) { if (cross) DQ_LAUNCH_LIFT(true, true); else DQ_LAUNCH_LIFT(true, false); }
        else { if (cross) DQ_LAUNCH_LIFT(false, true); else DQ_LAUNCH_LIFT(false, false); }
    } else {                                // large angles: straight-line (cos, sin) body
        if (pl->has_aj) { if (cross) DQ_LAUNCH(true, true); else DQ_LAUNCH(true, false); }
        else { if (cross) DQ_LAUNCH(false, true); else DQ_LAUNCH(false, false); }
    }
#undef DQ_LAUNCH_LIFT
#undef DQ_LAUNCH
    p->ctx->launches++;
    if (timed) {
        DQ_CUDA(cudaEventRecord(pl->ev[pl->ev_used + 1], st));
        pl->ev_used += 2;
    }
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

static bool rows_allow_scaled(const dq_ising* p, const double* rows, long long n_rows, double r) {
    double mx = fabs(atan(r));
    for (long long k = 0; k < n_rows; ++k)
        for (int q = 0; q < p->n; ++q) mx = std::max(mx, fabs(rows[k * p->row_len + 1 + p->n_zz + q]));
    return mx <= 1.0;              // |tan| <= 1.56: the scaled butterfly stays well conditioned
}

static bool in_j(const Plan* pl, int type, int pos) {
    for (int i = 0; i < 5; ++i)
        if (pl->jphys[type][i] == pos) return true;
    return false;
}

}  // namespace fused

// debug: copy the last launch's phase timestamps to the host (DQ_TRACE builds only)
extern "C" long long dq_debug_trace(dq_ising* p, long long* out, long long max_items) {
    fused::Plan* pl = fused::find_plan(p);
    if (!pl || !pl->trace.p) return 0;
    cudaStreamSynchronize(p->ctx->stream);
    long long n = std::min(max_items, pl->trace_items);
    cudaMemcpy(out, pl->trace.p, (size_t)n * 48 * sizeof(long long), cudaMemcpyDeviceToHost);
    return n;
}

// Sum of the event-timed durations of the pass-kernel launches of the last run (option time_launches).
int fused_launch_times(dq_ising* p, double* total_ms, double* n_launches) {
    fused::Plan* pl = fused::find_plan(p);
    *total_ms = 0.0;
    *n_launches = 0.0;
    if (!pl || pl->ev_used == 0) return DQ_OK;
    DQ_CUDA(cudaStreamSynchronize(p->ctx->stream));
    for (int i = 0; i + 1 < pl->ev_used; i += 2) {
        float ms = 0.f;
        DQ_CUDA(cudaEventElapsedTime(&ms, pl->ev[i], pl->ev[i + 1]));
        *total_ms += ms;
        *n_launches += 1.0;
    }
    return DQ_OK;
}

void fused_j_sets(int n, int* jl, int* jh) {
    const int a = 22 - n;
    for (int i = 0; i < 5; ++i) {
        jl[i] = fused::kJBits[0][i];
        const int t = fused::kJBits[1][i];
        jh[i] = t < a ? t : 10 + (t - a);
    }
}

int fused_supported(const dq_ising* p) {
    fused::Plan* pl = fused::get_plan(const_cast<dq_ising*>(p));
    return pl->ok ? 1 : 0;
}

void fused_release(dq_ising* p) {
    for (size_t i = 0; i < fused::g_plans.size(); ++i)
        if (fused::g_plans[i].first == p) {
            fused::Plan* pl = fused::g_plans[i].second;
            DevBuf* bufs[] = {&pl->d_types, &pl->jobs, &pl->steps, &pl->tc, &pl->kets, &pl->counters, &pl->partials,
                              &pl->out_index, &pl->work, &pl->rows, &pl->phi, &pl->uniform, &pl->afin, &pl->ea, &pl->ea_index,
                              &pl->d_maps};
            for (auto* b : bufs) b->release();
            for (cudaEvent_t e : pl->ev) cudaEventDestroy(e);
            delete pl;
            fused::g_plans.erase(fused::g_plans.begin() + i);
            return;
        }
}

// Evolve `batch` states in place through the same rows (API: dq_ising_evolve).
int fused_evolve(dq_ising* p, c128* d_states, int batch, const double* h_rows, int n_steps, double* d_energies,
                 bool want_states, const double* d_rows, int scaled_hint) {
    using namespace fused;
    Plan* pl = get_plan(p);
    DQ_REQUIRE(pl->ok, "fused engine unavailable for this problem");
    pl->ev_used = 0;
    cudaStream_t st = p->ctx->stream;
    // d_rows: the angle rows are already on the device (device-resident training); the caller vouches for the angle range
    const bool scaled = d_rows ? scaled_hint != 0 : rows_allow_scaled(p, h_rows, n_steps, 0.0);
    const size_t N = p->dim();
    const int tiles = 1 << pl->tiles_log2;
    DQ_TRY(pl->rows.reserve((size_t)std::max(1, n_steps) * p->row_len * sizeof(double)));
    if (n_steps)
        DQ_CUDA(cudaMemcpyAsync(pl->rows.p, d_rows ? d_rows : h_rows, (size_t)n_steps * p->row_len * sizeof(double),
                                d_rows ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    std::vector<SetupJob> jobs;
    Traj tr{0, n_steps, 0, !want_states, 0};       // energy-only: the final pass reduces instead of storing
    tr.scaled = scaled;
    add_traj(jobs, tr);
    DQ_TRY(run_setup(p, pl, jobs, pl->rows.as<double>()));
    std::vector<KetDesc> kets(batch);
    DQ_TRY(pl->partials.reserve((size_t)batch * tiles * sizeof(double)));
    DQ_TRY(maps_begin(p, pl, (size_t)batch));
    for (int g = 0; g < batch; ++g) {
        KetDesc& k = kets[g];
        k.src = d_states + (size_t)g * N;
        k.buf = d_states + (size_t)g * N;
        k.map_src = k.map_buf = map_index(pl, k.buf);
        DQ_REQUIRE(k.map_src >= 0, "fused engine: cuTensorMapEncodeTiled failed for a state buffer");
        k.steps = pl->steps.as<PassStep>();
        k.partial = pl->partials.as<double>() + (size_t)g * tiles;
        k.sigma = 0.0;
        k.escale = 1.0;
        k.n_pass = n_steps + 1;
        k.shift_kind = -1;
        k.sb0 = k.sb1 = 0;
        k.cls = 0;
        k.pad_ = 0;
        k.cross = nullptr;
        k.partial2 = nullptr;
        k.escale2 = 0.0;
    }
    DQ_TRY(maps_upload(p, pl));
    DQ_TRY(pl->kets.reserve(kets.size() * sizeof(KetDesc)));
    DQ_CUDA(cudaMemcpyAsync(pl->kets.p, kets.data(), kets.size() * sizeof(KetDesc), cudaMemcpyHostToDevice, st));
    const int G = auto_ket_group(p);
    for (int g0 = 0; g0 < batch; g0 += G)
        DQ_TRY(launch_group(p, pl, pl->kets.as<KetDesc>() + g0, std::min(G, batch - g0), n_steps + 1, scaled, 0.5));
    if (d_energies) {
        if (want_states) {
            DQ_TRY(gen_energy(p, d_states, batch, d_energies));
        } else {
            std::vector<int> idx(batch);
            for (int g = 0; g < batch; ++g) idx[g] = g;
            DQ_TRY(pl->out_index.reserve(batch * sizeof(int)));
            DQ_CUDA(cudaMemcpyAsync(pl->out_index.p, idx.data(), batch * sizeof(int), cudaMemcpyHostToDevice, st));
            k_sum_partials<<<batch, 32, 0, st>>>(pl->partials.as<double>(), tiles, tiles >> pl->sub_log2, pl->out_index.as<int>(), d_energies);
            p->ctx->launches++;
        }
    }
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

// Batched gradient samples from the staged tables (API: dq_ising_grad_run_staged).
int fused_grad_run(dq_ising* p) {
    using namespace fused;
    Plan* pl = get_plan(p);
    DQ_REQUIRE(pl->ok, "fused engine unavailable for this problem");
    pl->ev_used = 0;
    auto& s = p->st;
    cudaStream_t st = p->ctx->stream;
    const size_t N = p->dim();
    const int tiles = 1 << pl->tiles_log2;
    const int B = s.n_samples, n_shift = s.n_shift, kets_per = 2 * n_shift;
    const int G = auto_ket_group(p);
    // lifting form or straight-line form is decided per SAMPLE (its own angles), so a sample's result does not depend on
    // which other samples share its batch (bit-equal sharded runs, SURVEY 4 iv)
    auto scaled_of = [&](int b) { return s.scaled_sample.empty() ? s.scaled_ok : s.scaled_sample[b] != 0; };
    const bool linear = p->linear != 0;   // one shifted ket per term + the unshifted suffix state (see k_linear_fix)

    // ---- tables: rows_a (prefix) and rows_b (suffix) live in one device table ---------------------
    const long long np = s.prefix_off[B], ns = s.suffix_off[B];
    DQ_TRY(pl->rows.reserve((size_t)std::max<long long>(1, np + ns) * p->row_len * sizeof(double)));
    if (np) DQ_CUDA(cudaMemcpyAsync(pl->rows.p, p->rows_a.p, np * p->row_len * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (ns) DQ_CUDA(cudaMemcpyAsync(pl->rows.as<double>() + np * p->row_len, p->rows_b.p, ns * p->row_len * sizeof(double), cudaMemcpyDeviceToDevice, st));
    std::vector<SetupJob> jobs;
    std::vector<Traj> pre(B), sufL(B), sufH(B), sufA(B);
    for (int b = 0; b < B; ++b) {
        pre[b] = Traj{s.prefix_off[b], s.prefix_steps[b], 0, false, 0};
        sufL[b] = Traj{np + s.suffix_off[b], s.suffix_steps[b], 0, true, 0};
        sufH[b] = Traj{np + s.suffix_off[b], s.suffix_steps[b], 1, true, 0};
        pre[b].scaled = sufL[b].scaled = sufH[b].scaled = scaled_of(b);
        add_traj(jobs, pre[b]);
        add_traj(jobs, sufL[b]);
        add_traj(jobs, sufH[b]);
        if (linear) {
            sufA[b] = Traj{np + s.suffix_off[b], s.suffix_steps[b], 0, true, 0, true};
            sufA[b].scaled = scaled_of(b);
            add_traj(jobs, sufA[b]);
        }
    }
    DQ_TRY(run_setup(p, pl, jobs, pl->rows.as<double>()));

    // ---- buffers -------------------------------------------------------------------------------------
    DQ_TRY(pl->phi.reserve((size_t)B * N * sizeof(c128)));
    DQ_TRY(pl->work.reserve((size_t)G * N * sizeof(c128)));
    DQ_TRY(pl->partials.reserve((size_t)B * kets_per * tiles * sizeof(double)));
    if (linear) {
        DQ_TRY(pl->afin.reserve((size_t)B * N * sizeof(c128)));
        DQ_TRY(pl->ea.reserve(((size_t)B * tiles + B) * sizeof(double)));
    }
    const c128* psi0 = nullptr;
    if (s.uniform_psi0) {
        DQ_TRY(pl->uniform.reserve(N * sizeof(c128)));
        DQ_TRY(gen_fill_uniform(p, pl->uniform.as<c128>(), 1));
        psi0 = pl->uniform.as<c128>();
    } else {
        psi0 = s.psi0.as<c128>();
    }

    // ---- ket descriptors: B prefix kets, then per sample the shifted kets ordered by class ------------
    std::vector<KetDesc> kets;
    kets.reserve((size_t)B * (kets_per + 1));
    DQ_TRY(maps_begin(p, pl, (size_t)2 * B + G + 2));
    auto view_of = [&](const void* base) { return map_index(pl, base); };
    DQ_REQUIRE(view_of(psi0) >= 0, "fused engine: cuTensorMapEncodeTiled failed for a state buffer");
    for (int b = 0; b < B; ++b) {
        KetDesc k;
        k.src = psi0;
        k.buf = pl->phi.as<c128>() + (size_t)b * N;
        k.map_src = view_of(k.src);
        k.map_buf = view_of(k.buf);
        k.steps = pl->steps.as<PassStep>() + pre[b].step0;
        k.partial = nullptr;
        k.sigma = 0.0;
        k.escale = 1.0;
        k.n_pass = pre[b].n_steps + 1;
        k.shift_kind = -1;
        k.sb0 = k.sb1 = 0;
        k.cls = 0;
        k.pad_ = 0;
        k.cross = nullptr;
        k.partial2 = nullptr;
        k.escale2 = 0.0;
        kets.push_back(k);
    }
    struct Group { size_t first; int count; int max_pass; int chain; bool scaled; };   // chain > 0: groups of `chain` kets share a work ring
    std::vector<Group> groups;
    // launches are homogeneous in the butterfly form: consecutive samples of one form, at most G per launch
    auto group_runs = [&](size_t base) {
        for (int g0 = 0; g0 < B;) {
            int cnt = 1, mp = kets[base + g0].n_pass;
            while (g0 + cnt < B && cnt < G && scaled_of(g0 + cnt) == scaled_of(g0)) {
                mp = std::max(mp, kets[base + g0 + cnt].n_pass);
                ++cnt;
            }
            groups.push_back({base + g0, cnt, mp, 0, scaled_of(g0)});
            g0 += cnt;
        }
    };
    group_runs(0);
    if (linear) {                       // a_b = U(suffix of b) phi_b : stored (cross terms) and reduced (Ea)
        for (int b = 0; b < B; ++b) {
            KetDesc k = kets[b];
            k.src = pl->phi.as<c128>() + (size_t)b * N;
            k.buf = pl->afin.as<c128>() + (size_t)b * N;
            k.map_src = view_of(k.src);
            k.map_buf = view_of(k.buf);
            k.steps = pl->steps.as<PassStep>() + sufA[b].step0;
            k.partial = pl->ea.as<double>() + (size_t)b * tiles;
            k.n_pass = s.suffix_steps[b] + 1;
            kets.push_back(k);
        }
        group_runs((size_t)B);
    }
    // One launch per sample: its shifted kets (class 0 first, then class 1) form a chain of groups of G kets that share
    // the G work buffers -- ket k starts in the buffer of ket k - G as soon as that one is finished, so the SMs never
    // drain between groups.
    const int chain_len = std::max(G, (kMaxGroup / G) * G);
    for (int b = 0; b < B; ++b) {
        std::vector<KetDesc> mine;
        const double esc_x = scaled_of(b) ? 1.0 / (1.0 + s.r * s.r) : 1.0;
        for (int cls = 0; cls < 2; ++cls) {
            for (int i = 0; i < n_shift; ++i) {
                int kcls = 0, b0 = 0, b1 = 0;
                if (s.shift_kind[i] == 0) {
                    // a ZZ shift rides on the phase tables; start with the pass type in which the pair is not J-J
                    b0 = p->pa[s.shift_index[i]]; b1 = p->pb[s.shift_index[i]];
                    kcls = (in_j(pl, 0, b0) && in_j(pl, 0, b1)) ? 1 : 0;
                    if (kcls == 1 && in_j(pl, 1, b0) && in_j(pl, 1, b1)) {
                        set_error("fused engine: ZZ shift pair (%d,%d) is J-J in both pass types", b0, b1);
                        return DQ_ERR_UNSUPPORTED;
                    }
                } else {
                    // an X shift is one more butterfly: start with the pass type that rotates that bit
                    b0 = p->bitpos[s.shift_index[i]];
                    kcls = b0 >= 10 ? 1 : 0;
                }
                if (kcls != cls) continue;
                for (int sg = 0; sg < (linear ? 1 : 2); ++sg) {
                    KetDesc k;
                    k.cross = linear ? pl->afin.as<c128>() + (size_t)b * N : nullptr;
                    k.partial2 = linear ? pl->partials.as<double>() + ((size_t)b * kets_per + 2 * i + 1) * tiles : nullptr;
                    k.escale2 = s.shift_kind[i] == 1 ? sqrt(esc_x) : 1.0;
                    k.src = pl->phi.as<c128>() + (size_t)b * N;
                    k.map_src = view_of(k.src);
                    k.map_buf = -1;
                    k.buf = nullptr;                // slot assigned below
                    k.steps = pl->steps.as<PassStep>() + (cls == 0 ? sufL[b].step0 : sufH[b].step0);
                    const size_t kidx = (size_t)b * kets_per + 2 * i + sg;
                    k.partial = pl->partials.as<double>() + kidx * tiles;
                    k.sigma = sg == 0 ? +1.0 : -1.0;
                    k.escale = s.shift_kind[i] == 1 ? esc_x : 1.0;
                    k.n_pass = s.suffix_steps[b] + 1;
                    k.shift_kind = s.shift_kind[i];
                    k.sb0 = b0;
                    k.sb1 = b1;
                    k.cls = cls;
                    k.pad_ = 0;
                    mine.push_back(k);
                }
            }
        }
        for (size_t g0 = 0; g0 < mine.size(); g0 += chain_len) {
            const int cnt = (int)std::min<size_t>(chain_len, mine.size() - g0);
            for (int g = 0; g < cnt; ++g) {
                mine[g0 + g].buf = pl->work.as<c128>() + (size_t)(g % G) * N;
                mine[g0 + g].map_buf = view_of(mine[g0 + g].buf);
            }
            groups.push_back({kets.size() + g0, cnt, s.suffix_steps[b] + 1, G, scaled_of(b)});
        }
        kets.insert(kets.end(), mine.begin(), mine.end());
    }
    std::vector<int> out_index((size_t)B * kets_per);
    for (size_t i = 0; i < out_index.size(); ++i) out_index[i] = (int)i;
    for (const KetDesc& k : kets)
        DQ_REQUIRE(k.map_src >= 0 && k.map_buf >= 0, "fused engine: cuTensorMapEncodeTiled failed for a state buffer");
    DQ_TRY(maps_upload(p, pl));
    DQ_TRY(pl->kets.reserve(kets.size() * sizeof(KetDesc)));
    DQ_CUDA(cudaMemcpyAsync(pl->kets.p, kets.data(), kets.size() * sizeof(KetDesc), cudaMemcpyHostToDevice, st));
    DQ_TRY(pl->out_index.reserve(out_index.size() * sizeof(int)));
    DQ_CUDA(cudaMemcpyAsync(pl->out_index.p, out_index.data(), out_index.size() * sizeof(int), cudaMemcpyHostToDevice, st));

    for (const Group& g : groups)
        DQ_TRY(launch_group(p, pl, pl->kets.as<KetDesc>() + g.first, g.count, g.max_pass, g.scaled, s.r, g.chain));

    k_sum_partials<<<B * kets_per, 32, 0, st>>>(pl->partials.as<double>(), tiles, tiles >> pl->sub_log2, pl->out_index.as<int>(),
                                                 p->energies.as<double>());
    p->ctx->launches++;
    if (linear) {
        double* ea_out = pl->ea.as<double>() + (size_t)B * tiles;
        k_sum_partials<<<B, 32, 0, st>>>(pl->ea.as<double>(), tiles, tiles >> pl->sub_log2, pl->out_index.as<int>(), ea_out);
        const int total = B * kets_per;
        k_linear_fix<<<(total + 255) / 256, 256, 0, st>>>(p->energies.as<double>(), ea_out, kets_per, total, s.r);
        p->ctx->launches += 2;
    }
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

}  // namespace dq
#endif  // DQ_PROBE_ONLY
