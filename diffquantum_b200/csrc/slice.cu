// Slice kernels for a single state distributed over ranks on its high-order index bits (SURVEY 8e,
// BASELINE configs[4]).  A rank owns 2^L consecutive amplitudes; the global basis index of local x is
// (high_bits << L) | x.  The diagonal phase needs no communication in ANY layout — the host passes the
// current physical bit position of every ZZ endpoint — and an X rotation is local whenever its qubit
// currently sits on a bit below L.  Qubits on the top g bits are made local by an all-to-all that swaps
// bits [L-g, L) with the rank bits (done by the host through NCCL, diffquantum_b200/distributed.py).
// Step semantics: the per-term product of diffqc.cc:155-164.
#include <algorithm>
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <utility>
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxPairs = 256;
constexpr int kMaxPeers = 16;
constexpr int kTileBits12 = 12;                     // the TMA tile kernel: always 2^12 amplitudes (64 KiB)
constexpr unsigned kRingSlots = 256;                // device argument tables in flight (ctx->slice_ring)

struct PhaseArgs {
    int n_zz;
    unsigned char a[kMaxPairs], b[kMaxPairs];
    double ang[kMaxPairs];
    double c0;
};

__global__ void __launch_bounds__(kThreads) k_slice_phase(double2* __restrict__ psi, int L, unsigned long long high,
                                                          const PhaseArgs* __restrict__ pa) {
    __shared__ double ang[kMaxPairs];
    __shared__ unsigned char ba[kMaxPairs], bb[kMaxPairs];
    const int n_zz = pa->n_zz;
    for (int e = threadIdx.x; e < n_zz; e += blockDim.x) { ang[e] = pa->ang[e]; ba[e] = pa->a[e]; bb[e] = pa->b[e]; }
    __syncthreads();
    const double c0 = pa->c0;
    const size_t N = (size_t)1 << L;
    const unsigned long long hi = high << L;
    for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < N; x += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long g = hi | x;
        double a = c0;
        for (int e = 0; e < n_zz; ++e) a += (((g >> ba[e]) ^ (g >> bb[e])) & 1ull) ? -ang[e] : ang[e];
        double sn, cs;
        sincos(a, &sn, &cs);
        const double2 v = psi[x];
        psi[x] = make_double2(v.x * cs + v.y * sn, v.y * cs - v.x * sn);
    }
}

// Diagonal phase with incremental angles: a thread owns one value of the low S = L - B index bits and walks the top B
// local bits in Gray-code order, so consecutive amplitudes differ in ONE bit p and
//     exp(-i a(g ^ 1<<p)) = exp(-i a(g)) * prod_{pairs e containing p} (cos 2w_e + i sigma_e(g) sin 2w_e),   sigma_e = z_a z_b.
// One full angle evaluation + sincos per 2^B amplitudes instead of one per amplitude (the full form is compute-bound:
// ~400 instructions per amplitude at 45 pairs); accesses stay coalesced because lanes differ in the low bits.
constexpr int kGrayBits = 6, kGrayDeg = 16;
struct GrayArgs {
    int B, S;
    int deg[kGrayBits];
    unsigned char other[kGrayBits][kGrayDeg];
    double c2[kGrayBits][kGrayDeg], s2[kGrayBits][kGrayDeg];
};

__global__ void __launch_bounds__(kThreads) k_slice_phase_gray(double2* __restrict__ psi, int L, unsigned long long high,
                                                               const PhaseArgs* __restrict__ pa, const GrayArgs* __restrict__ ga) {
    __shared__ double ang[kMaxPairs];
    __shared__ unsigned char ba[kMaxPairs], bb[kMaxPairs];
    __shared__ GrayArgs G;
    const int n_zz = pa->n_zz;
    for (int e = threadIdx.x; e < n_zz; e += blockDim.x) { ang[e] = pa->ang[e]; ba[e] = pa->a[e]; bb[e] = pa->b[e]; }
    if (threadIdx.x == 0) G = *ga;
    __syncthreads();
    const double c0 = pa->c0;
    const int B = G.B, S = G.S;
    const size_t n_low = (size_t)1 << S;
    const unsigned long long hi = high << L;
    for (size_t low = blockIdx.x * (size_t)blockDim.x + threadIdx.x; low < n_low; low += (size_t)gridDim.x * blockDim.x) {
        unsigned long long g = hi | low;
        double a = c0;
        for (int e = 0; e < n_zz; ++e) a += (((g >> ba[e]) ^ (g >> bb[e])) & 1ull) ? -ang[e] : ang[e];
        double sn, cs;
        sincos(a, &sn, &cs);
        double2 ph = make_double2(cs, -sn);                                  // exp(-i a)
        {
            const double2 v = psi[low];
            psi[low] = make_double2(v.x * ph.x - v.y * ph.y, v.x * ph.y + v.y * ph.x);
        }
        for (int k = 1; k < (1 << B); ++k) {
            const int j = __ffs(k) - 1, p = S + j;
            const int d = G.deg[j];
            for (int q = 0; q < d; ++q) {
                const bool differ = (((g >> G.other[j][q]) ^ (g >> p)) & 1ull) != 0;
                const double c2 = G.c2[j][q], s2 = differ ? -G.s2[j][q] : G.s2[j][q];
                ph = make_double2(ph.x * c2 - ph.y * s2, ph.x * s2 + ph.y * c2);
            }
            g ^= 1ull << p;
            const size_t x = (size_t)(g & ((1ull << L) - 1ull));
            const double2 v = psi[x];
            psi[x] = make_double2(v.x * ph.x - v.y * ph.y, v.x * ph.y + v.y * ph.x);
        }
    }
}

__global__ void __launch_bounds__(kThreads) k_slice_rx(double2* __restrict__ psi, int L, int bit, double c, double s) {
    const size_t half = (size_t)1 << (L - 1);
    const size_t low = ((size_t)1 << bit) - 1;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
        const size_t x0 = ((i & ~low) << 1) | (i & low), x1 = x0 | ((size_t)1 << bit);
        const double2 a = psi[x0], b = psi[x1];
        psi[x0] = make_double2(c * a.x + s * b.y, c * a.y - s * b.x);
        psi[x1] = make_double2(c * b.x + s * a.y, c * b.y - s * a.x);
    }
}

// ---- fused X-rotation pass ---------------------------------------------------------------------------------------
// X rotations on different qubits commute, so all rotations of a step whose bits fit one TILE are applied in a single
// read + write of the slice.  A tile is 2^T amplitudes (T <= 12, 64 KiB of shared memory) spanned by T physical bits
// pos[0] < pos[1] < ...: the `lo` lowest are bits 0 .. lo-1 (contiguous 16 * 2^lo byte runs in global memory), the others
// are the rotation targets of this pass.  Inside the tile two bits are rotated per shared-memory round trip (radix 4).
struct TileArgs {
    int T, lo;
    int n_active;
    unsigned char pos[12];          // physical bit of tile bit i
    unsigned char active[12];       // tile bits to rotate
    double c[12], s[12];            // per active entry; scaled form: c = 1, s = tan(theta)
    unsigned long long mask;        // OR of 1 << pos[i]
    double post;                    // scaled form: product of the cosines, applied once when the tile is stored
    int scaled;
    // exchange fused into the stores of this pass (the last local pass before the global<->local qubit swap): element x of
    // the slice goes to rank j = x >> (L - g) at peer[j][(rank << (L - g)) | (x & (2^(L-g) - 1))] -- the all-to-all of
    // diffquantum_b200/distributed.py written straight into the peers' receive buffers over NVLink.  scatter_g = 0: in place.
    int scatter_g, scatter_rank;
    double2* peer[kMaxPeers];
    // rotations applied BEFORE the phase (k_slice_rx_tma with PHASE only): the rotations of the PREVIOUS step on this tile's
    // bits, so that a pass reads  [rotations of step p] [phase of step p+1] [rotations of step p+1]  and a step costs one
    // pass less; scaled form: their cosines are part of `post`
    int n_pre;
    unsigned char pre_active[12];
    double pre_c[12], pre_s[12];
};

// exp(-i theta X) on a pair.  SCALED: a' = a - i tan(theta) b (one FMA per real component, the cosines are applied once per
// pass) - the rotation passes are co-limited by the FP64 pipe, so halving its work is worth a template parameter.
template <bool SCALED>
__device__ __forceinline__ void rot_pair(double2& a, double2& b, double c, double s) {
    const double2 a0 = a, b0 = b;
    if (SCALED) {
        a = make_double2(fma(s, b0.y, a0.x), fma(-s, b0.x, a0.y));
        b = make_double2(fma(s, a0.y, b0.x), fma(-s, a0.x, b0.y));
    } else {
        a = make_double2(fma(s, b0.y, c * a0.x), fma(-s, b0.x, c * a0.y));
        b = make_double2(fma(s, a0.y, c * b0.x), fma(-s, a0.x, c * b0.y));
    }
}

// shared-memory slot of tile element e: XOR swizzle of the 16-byte unit inside each 128-byte row, so that the low-bit
// rounds (threads 32 / 64 / 128 bytes apart) stay bank-conflict free
__device__ __forceinline__ int tslot(int e) { return e ^ ((e >> 3) & 7); }

// The rotation rounds of one tile: every active bit, three (two, one) per shared-memory round trip.  `tid` < kThreads is the
// thread's index in its team of kThreads threads, `bar` the named barrier the team synchronises on (0 = the whole CTA when the
// CTA is one team).  pre != 1: the amplitudes are multiplied by `pre` as they are first read (scaled form: the product of the
// cosines, for callers that cannot apply it on the way out).
__device__ __forceinline__ void team_bar(int bar) { asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(kThreads) : "memory"); }

template <bool SCALED>
__device__ __forceinline__ void tile_rounds(double2* __restrict__ tile, const int T, const int n_active, const unsigned char* active,
                                            const double* rc, const double* rs, const int tid, const int bar, const double pre) {
    const int n_el = 1 << T;
    bool first = pre != 1.0;
    int k = 0;
    for (; k + 2 < n_active; k += 3) {                  // three bits per round trip (`active` is ascending)
        const int i0 = active[k], i1 = active[k + 1], i2 = active[k + 2];
        const double c0 = rc[k], s0 = rs[k], c1 = rc[k + 1], s1 = rs[k + 1], c2 = rc[k + 2], s2 = rs[k + 2];
        for (int q = tid; q < (n_el >> 3); q += kThreads) {
            int e = ((q >> i0) << (i0 + 1)) | (q & ((1 << i0) - 1));
            e = ((e >> i1) << (i1 + 1)) | (e & ((1 << i1) - 1));
            e = ((e >> i2) << (i2 + 1)) | (e & ((1 << i2) - 1));
            int sl[8];
            double2 v[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                sl[b] = tslot(e | ((b & 1) << i0) | (((b >> 1) & 1) << i1) | (((b >> 2) & 1) << i2));
                v[b] = tile[sl[b]];
            }
            if (first) {
#pragma unroll
                for (int b = 0; b < 8; ++b) v[b] = make_double2(v[b].x * pre, v[b].y * pre);
            }
            rot_pair<SCALED>(v[0], v[1], c0, s0); rot_pair<SCALED>(v[2], v[3], c0, s0);
            rot_pair<SCALED>(v[4], v[5], c0, s0); rot_pair<SCALED>(v[6], v[7], c0, s0);
            rot_pair<SCALED>(v[0], v[2], c1, s1); rot_pair<SCALED>(v[1], v[3], c1, s1);
            rot_pair<SCALED>(v[4], v[6], c1, s1); rot_pair<SCALED>(v[5], v[7], c1, s1);
            rot_pair<SCALED>(v[0], v[4], c2, s2); rot_pair<SCALED>(v[1], v[5], c2, s2);
            rot_pair<SCALED>(v[2], v[6], c2, s2); rot_pair<SCALED>(v[3], v[7], c2, s2);
#pragma unroll
            for (int b = 0; b < 8; ++b) tile[sl[b]] = v[b];
        }
        first = false;
        team_bar(bar);
    }
    for (; k + 1 < n_active; k += 2) {                  // two bits per round trip
        const int i0 = min(active[k], active[k + 1]), i1 = max(active[k], active[k + 1]);
        const double c0 = active[k] == i0 ? rc[k] : rc[k + 1], s0 = active[k] == i0 ? rs[k] : rs[k + 1];
        const double c1 = active[k] == i0 ? rc[k + 1] : rc[k], s1 = active[k] == i0 ? rs[k + 1] : rs[k];
        for (int q = tid; q < (n_el >> 2); q += kThreads) {
            int e = ((q >> i0) << (i0 + 1)) | (q & ((1 << i0) - 1));            // zero at bit i0
            e = ((e >> i1) << (i1 + 1)) | (e & ((1 << i1) - 1));                // zero at bit i1
            const int s00 = tslot(e), s01 = tslot(e | (1 << i0)), s10 = tslot(e | (1 << i1)), s11 = tslot(e | (1 << i0) | (1 << i1));
            double2 v00 = tile[s00], v01 = tile[s01], v10 = tile[s10], v11 = tile[s11];
            if (first) {
                v00 = make_double2(v00.x * pre, v00.y * pre); v01 = make_double2(v01.x * pre, v01.y * pre);
                v10 = make_double2(v10.x * pre, v10.y * pre); v11 = make_double2(v11.x * pre, v11.y * pre);
            }
            rot_pair<SCALED>(v00, v01, c0, s0);
            rot_pair<SCALED>(v10, v11, c0, s0);
            rot_pair<SCALED>(v00, v10, c1, s1);
            rot_pair<SCALED>(v01, v11, c1, s1);
            tile[s00] = v00; tile[s01] = v01; tile[s10] = v10; tile[s11] = v11;
        }
        first = false;
        team_bar(bar);
    }
    if (k < n_active) {
        const int i0 = active[k];
        const double c0 = rc[k], s0 = rs[k];
        for (int q = tid; q < (n_el >> 1); q += kThreads) {
            const int e = ((q >> i0) << (i0 + 1)) | (q & ((1 << i0) - 1));
            const int sa = tslot(e), sb = tslot(e | (1 << i0));
            double2 a = tile[sa], b = tile[sb];
            if (first) { a = make_double2(a.x * pre, a.y * pre); b = make_double2(b.x * pre, b.y * pre); }
            rot_pair<SCALED>(a, b, c0, s0);
            tile[sa] = a; tile[sb] = b;
        }
        first = false;
        team_bar(bar);
    }
}

// PHASE (contiguous 12-bit tiles only): the diagonal phase of the SAME step rides on this pass -- applied to the tile in shared
// memory right after it has landed, before the first rotation round -- so a product-formula step costs one pass over the
// slice less.  Thread t of a team owns tile bits 0..7 (= its index) and walks tile bits 8..11 in Gray-code order like
// k_slice_phase_gray: one angle evaluation + sincos per 16 amplitudes, then ONE complex multiply per flipped bit.
constexpr int kWalkBits = 4, kWalkLo = 8, kTileBits = 12, kMaxIn = 96;
// Tables built once per CTA from the pair list (tile bit = physical bit k < 12 in a contiguous pass):
//   out_*   per tile bit k: its pairs whose other end lies OUTSIDE the tile -> per tile a field h[k] on z_k
//   in_*    pairs with both ends inside the tile
//   oo_*    pairs with both ends outside the tile -> per tile a constant
//   w_*     per walk bit j (tile bit 8 + j): exp(2i w) factors of its pairs with thread bits (w_in) and with outside bits (w_out,
//           per tile folded into ONE factor f_out[j]; z_p = -1 takes its conjugate)
//   ww      pairs among the walk bits themselves: one factor per value of the walk bits, the same for every thread and tile
struct PhaseTabs {
    int out_n[kTileBits], w_in_n[kWalkBits], w_out_n[kWalkBits], in_n, oo_n, has_ww;
    unsigned char out_other[kTileBits][kGrayDeg], in_a[kMaxIn], in_b[kMaxIn], oo_a[kMaxPairs], oo_b[kMaxPairs];
    unsigned char w_in_other[kWalkBits][kGrayDeg], w_out_other[kWalkBits][kGrayDeg];
    double out_ang[kTileBits][kGrayDeg], in_ang[kMaxIn], oo_ang[kMaxPairs], c0;
    double w_in_c2[kWalkBits][kGrayDeg], w_in_s2[kWalkBits][kGrayDeg], w_out_c2[kWalkBits][kGrayDeg], w_out_s2[kWalkBits][kGrayDeg];
    double2 ww[1 << kWalkBits];
};
struct PhaseTile {                   // per tile (one per team): fields of the outside bits, the outside-outside constant, outside factors
    double h_field[kTileBits], c_tile;
    double2 f_out[kWalkBits];
};

// tile bit of physical bit `phys` (-1: the tile does not span it)
__device__ __forceinline__ int tile_bit_of(const TileArgs& A, const int phys) {
    int r = -1;
    for (int i = 0; i < A.T; ++i) r = A.pos[i] == phys ? i : r;
    return r;
}

// once per CTA; t = 0 .. (at least 48) distinct threads; a barrier must follow.  has_ww must be 0 on entry.  The tile (A.T = 12)
// may span any physical bits: pair ends are classified by tile membership, in-tile ends are stored as TILE bits, outside ends
// as PHYSICAL bits (they are read from the global index of the tile's element 0).
__device__ __forceinline__ void phase_setup(PhaseTabs& P, const PhaseArgs* __restrict__ pa, const TileArgs& A, const int t) {
    const int n_zz = pa->n_zz;
    if (t < kTileBits) {
        const int k = t, pk = A.pos[k];
        int d = 0;
        for (int e = 0; e < n_zz; ++e) {
            const int a = pa->a[e], b = pa->b[e];
            if (a != pk && b != pk) continue;
            const int o = a == pk ? b : a;
            if (tile_bit_of(A, o) >= 0) continue;
            P.out_other[k][d] = (unsigned char)o;
            P.out_ang[k][d] = pa->ang[e];
            ++d;
        }
        P.out_n[k] = d;
    } else if (t == kTileBits) {
        int d = 0, o = 0;
        for (int e = 0; e < n_zz; ++e) {
            const int a = pa->a[e], b = pa->b[e];
            const int ta = tile_bit_of(A, a), tb = tile_bit_of(A, b);
            if (ta >= 0 && tb >= 0) { P.in_a[d] = (unsigned char)ta; P.in_b[d] = (unsigned char)tb; P.in_ang[d] = pa->ang[e]; ++d; }
            else if (ta < 0 && tb < 0) { P.oo_a[o] = (unsigned char)a; P.oo_b[o] = (unsigned char)b; P.oo_ang[o] = pa->ang[e]; ++o; }
        }
        P.in_n = d;
        P.oo_n = o;
        P.c0 = pa->c0;
    } else if (t < kTileBits + 1 + kWalkBits) {
        const int j = t - kTileBits - 1, p = A.pos[kWalkLo + j];
        int di = 0, dout = 0;
        for (int e = 0; e < n_zz; ++e) {
            const int a = pa->a[e], b = pa->b[e];
            if (a != p && b != p) continue;
            const int o = a == p ? b : a, to = tile_bit_of(A, o);
            if (to >= kWalkLo) continue;                                 // walk-walk pair: ww[] below
            double sn, cs;
            sincos(2.0 * pa->ang[e], &sn, &cs);
            if (to >= 0) { P.w_in_other[j][di] = (unsigned char)to; P.w_in_c2[j][di] = cs; P.w_in_s2[j][di] = sn; ++di; }
            else { P.w_out_other[j][dout] = (unsigned char)o; P.w_out_c2[j][dout] = cs; P.w_out_s2[j][dout] = sn; ++dout; }
        }
        P.w_in_n[j] = di;
        P.w_out_n[j] = dout;
    } else if (t >= 32 && t < 32 + (1 << kWalkBits)) {
        // pairs among the walk bits themselves: relative to walk value 0 a pair whose two bits differ contributes exp(+2i w)
        const int w = t - 32;
        double2 f = make_double2(1.0, 0.0);
        for (int e = 0; e < n_zz; ++e) {
            const int ta = tile_bit_of(A, pa->a[e]), tb = tile_bit_of(A, pa->b[e]);
            if (ta < kWalkLo || tb < kWalkLo) continue;
            P.has_ww = 1;
            if ((((w >> (ta - kWalkLo)) ^ (w >> (tb - kWalkLo))) & 1) == 0) continue;
            double sn, cs;
            sincos(2.0 * pa->ang[e], &sn, &cs);
            f = make_double2(f.x * cs - f.y * sn, f.x * sn + f.y * cs);
        }
        P.ww[w] = f;
    }
}

// per tile; threads t < 16 and 32 <= t < 64 of the team do the work (fixed summation / product order); G = global index of the
// tile's element 0.  A team barrier must follow.  The constant over the outside-outside pairs is a warp-wide sum (lane q takes
// pairs q, q + 32, ...; a serial loop over ~20 pairs was ~2000 cycles of dependent latency per tile with the team waiting).
__device__ __forceinline__ void phase_tile_setup(const PhaseTabs& P, PhaseTile& Q, const unsigned long long G, const int t) {
    if (t < kTileBits) {
        const int k = t;
        double hsum = 0.0;
        for (int q = 0; q < P.out_n[k]; ++q) hsum += ((G >> P.out_other[k][q]) & 1ull) ? -P.out_ang[k][q] : P.out_ang[k][q];
        Q.h_field[k] = hsum;
    } else if (t < kTileBits + kWalkBits) {
        const int j = t - kTileBits;
        double2 f = make_double2(1.0, 0.0);
        for (int q = 0; q < P.w_out_n[j]; ++q) {
            const double c2 = P.w_out_c2[j][q], s2 = ((G >> P.w_out_other[j][q]) & 1ull) ? -P.w_out_s2[j][q] : P.w_out_s2[j][q];
            f = make_double2(f.x * c2 - f.y * s2, f.x * s2 + f.y * c2);
        }
        Q.f_out[j] = f;
    } else if (t >= 32 && t < 64) {                                  // one whole warp
        const int lane = t - 32;
        double c = 0.0;
        for (int q = lane; q < P.oo_n; q += 32) c += (((G >> P.oo_a[q]) ^ (G >> P.oo_b[q])) & 1ull) ? -P.oo_ang[q] : P.oo_ang[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) Q.c_tile = P.c0 + c;
    }
}

// what a thread's part of the phase does not owe to the tile: the angle of the pairs inside the tile at walk bits 0 and the
// flip factors of the walk bits from their pairs with the thread's own bits -- once per launch, not once per tile
struct PhaseThread {
    double a_in;
    double2 f_in[kWalkBits];
};
__device__ __forceinline__ void phase_thread_init(const PhaseTabs& P, PhaseThread& H, const int t) {
    double a = 0.0;
    for (int q = 0; q < P.in_n; ++q) a += (((t >> P.in_a[q]) ^ (t >> P.in_b[q])) & 1) ? -P.in_ang[q] : P.in_ang[q];
    H.a_in = a;
#pragma unroll
    for (int j = 0; j < kWalkBits; ++j) {
        double2 f = make_double2(1.0, 0.0);
        for (int q = 0; q < P.w_in_n[j]; ++q) {
            const double c2 = P.w_in_c2[j][q], s2 = ((t >> P.w_in_other[j][q]) & 1) ? -P.w_in_s2[j][q] : P.w_in_s2[j][q];
            f = make_double2(f.x * c2 - f.y * s2, f.x * s2 + f.y * c2);
        }
        H.f_in[j] = f;
    }
}

// per tile, every thread of the team (t = 0..255 = tile bits 0..7); the tile must have landed, Q must be complete.
// The 16 phases of a thread (walk bits 8..11) are a doubling product ph(w) = ph(0) * prod_{j in w} F[j] * ww[w], evaluated depth
// first (four live partial products): dependency depth 4 instead of the 15-step chain of a Gray-code walk, which left the FP64
// pipe waiting on its own latency (the phase cost as much as five rotation rounds).
__device__ __forceinline__ double2 cmul2(const double2 a, const double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ void phase_apply(const PhaseTabs& P, const PhaseTile& Q, const PhaseThread& H, double2* __restrict__ tile,
                                            const int t, const double scale = 1.0) {
    double hs[kTileBits];
#pragma unroll
    for (int k = 0; k < kTileBits; ++k) hs[k] = ((t >> k) & 1) ? -Q.h_field[k] : Q.h_field[k];     // walk bits are 0 in t
    const double a = ((Q.c_tile + H.a_in) + ((hs[0] + hs[1]) + (hs[2] + hs[3]))) + (((hs[4] + hs[5]) + (hs[6] + hs[7])) + ((hs[8] + hs[9]) + (hs[10] + hs[11])));
    double sn, cs;
    sincos(a, &sn, &cs);
    const double2 ph = make_double2(cs * scale, -sn * scale);        // scale * exp(-i a)
    // flip factor of walk bit j for THIS thread and tile (z_j: +1 -> -1, the other walk bits at +1): the outside pairs (f_out,
    // per tile) times the pairs with this thread's own bits (f_in, per launch)
    double2 F[kWalkBits];
#pragma unroll
    for (int j = 0; j < kWalkBits; ++j) F[j] = cmul2(Q.f_out[j], H.f_in[j]);
    const bool any_ww = P.has_ww != 0;
#pragma unroll
    for (int b3 = 0; b3 < 2; ++b3) {
        const double2 p3 = b3 ? cmul2(ph, F[3]) : ph;
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2) {
            const double2 p2 = b2 ? cmul2(p3, F[2]) : p3;
#pragma unroll
            for (int b1 = 0; b1 < 2; ++b1) {
                const double2 p1 = b1 ? cmul2(p2, F[1]) : p2;
#pragma unroll
                for (int b0 = 0; b0 < 2; ++b0) {
                    const int w = (b3 << 3) | (b2 << 2) | (b1 << 1) | b0;
                    double2 q = b0 ? cmul2(p1, F[0]) : p1;
                    if (any_ww) q = cmul2(q, P.ww[w]);
                    const int sl = tslot(t | (w << kWalkLo));
                    tile[sl] = cmul2(tile[sl], q);
                }
            }
        }
    }
}

template <bool CONTIG, bool SCALED, bool PHASE>
__global__ void __launch_bounds__(kThreads, 3) k_slice_rx_tile(double2* __restrict__ psi, int L, const TileArgs* __restrict__ ta,
                                                               const PhaseArgs* __restrict__ pa, unsigned long long high) {
    extern __shared__ __align__(16) double2 tile[];
    __shared__ TileArgs A;
    __shared__ PhaseTabs PT[1];              // ~8 KB; unused (and removed by the compiler) without PHASE
    __shared__ PhaseTile PQ;
    if (threadIdx.x == 0) { A = *ta; PT[0].has_ww = 0; }
    __syncthreads();
    if (PHASE) phase_setup(PT[0], pa, A, threadIdx.x);
    __syncthreads();
    PhaseThread PH;
    if (PHASE) phase_thread_init(PT[0], PH, threadIdx.x);
    const int T = A.T, lo = A.lo, n_el = 1 << T;
    const unsigned lowmask = (1u << lo) - 1u;
    const unsigned long long n_tiles = 1ull << (L - T);
    // physical offset of the non-contiguous tile bits, tabulated once per CTA (at most 2^9 entries: lo >= 3 when T = 12)
    __shared__ unsigned long long hi_off[512];
    for (int v = threadIdx.x; v < (1 << (T - lo)); v += kThreads) {
        unsigned long long off = 0;
        for (int i = lo; i < T; ++i) off |= (unsigned long long)((v >> (i - lo)) & 1) << A.pos[i];
        hi_off[v] = off;
    }
    __syncthreads();
    for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        // deposit the bits of t into the physical positions the tile does not span.  A scatter pass fills the peer-selecting
        // positions [L - g, L) FIRST and starts at its own rank: consecutive tiles (= the CTAs running at any moment) go to
        // different peers and no two ranks aim at the same peer with the same tile index -- with the plain order every CTA of
        // every rank wrote to peer 0 first, then all to peer 1, ... and each receiver's NVLink ingress took the traffic of all
        // senders in turn (measured at n = 32 on 8 GPUs: 36 ms per step instead of 21)
        unsigned long long base = 0;
        {
            unsigned long long rest = t;
            const int p_split = A.scatter_g ? L - A.scatter_g : L;
            if (A.scatter_g) {
                int free_top = 0;
                for (int p = p_split; p < L; ++p) free_top += ((A.mask >> p) & 1ull) ? 0 : 1;
                rest = (rest & ~((1ull << free_top) - 1ull)) | ((rest + (unsigned long long)A.scatter_rank) & ((1ull << free_top) - 1ull));
                for (int p = p_split; p < L; ++p)
                    if (!((A.mask >> p) & 1ull)) { base |= (rest & 1ull) << p; rest >>= 1; }
            }
            for (int p = 0; p < p_split; ++p)
                if (!((A.mask >> p) & 1ull)) { base |= (rest & 1ull) << p; rest >>= 1; }
        }
        // asynchronous global -> shared copies (LDGSTS): all of a thread's 16-byte pieces are in flight at once and never
        // pass through registers (with plain loads the first STS of each group waited for its LDG: 60 % of the stall samples)
        for (int e = threadIdx.x; e < n_el; e += kThreads) {
            const double2* src = psi + (base + (CONTIG ? (unsigned long long)e : ((e & lowmask) | hi_off[e >> lo])));
            const unsigned dst = (unsigned)__cvta_generic_to_shared(tile + tslot(e));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (PHASE) phase_tile_setup(PT[0], PQ, (high << L) | base, threadIdx.x);     // while the tile is in flight
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (PHASE) {
            phase_apply(PT[0], PQ, PH, tile, threadIdx.x);
            __syncthreads();
        }
        tile_rounds<SCALED>(tile, A.T, A.n_active, A.active, A.c, A.s, threadIdx.x, 0, 1.0);
        for (int e = threadIdx.x; e < n_el; e += kThreads)
        {
            double2 v = tile[tslot(e)];
            if (SCALED) v = make_double2(v.x * A.post, v.y * A.post);
            const unsigned long long x = base + (CONTIG ? (unsigned long long)e : ((e & lowmask) | hi_off[e >> lo]));
            if (A.scatter_g) {
                const int sh = L - A.scatter_g;
                A.peer[x >> sh][((unsigned long long)A.scatter_rank << sh) | (x & ((1ull << sh) - 1ull))] = v;
            } else {
                psi[x] = v;
            }
        }
        __syncthreads();
    }
}

// ---- the same pass with TMA tiles, three buffers per SM --------------------------------------------------------------
// One persistent CTA per SM: two consumer teams of kThreads threads, a loader warp and a storer warp (the structure of
// k_fused_ws, without dependencies between tiles).  The tile arrives by ONE tensor load (cp.async.bulk.tensor, hardware
// 128-byte swizzle = tslot, mbarrier completion) and leaves by ONE tensor store; while a team rotates tile k the loader already
// has tile k + 1 (the other team's) and k + 2 in flight and the storer drains tile k - 1, so neither the load latency nor
// the write-back sits on a team's critical path (k_slice_rx_tile: load, rotate and store of a tile are serial in its CTA, and
// the three resident CTAs only partly cover for one another).  The tile is described by a tensor of rank <= 5 over the slice:
// dimension 0 = 8 amplitudes (128 bytes), every further dimension one run of consecutive tile bits, spanning up to the next
// run so that its coordinate also carries the tile-index bits lying in between (box = the run).  Passes whose tile bits form
// more than four runs, scatter passes and slices below 2^12 stay on k_slice_rx_tile.
constexpr int kTmaBufs = 3, kTmaTeams = 2, kTmaThreads = kTmaTeams * kThreads + 64;
struct TmaGeom {
    int rank;                        // 2..5
    int start[5], nbits[5];          // coordinate of dimension d: (base >> start[d]) & (2^nbits[d] - 1) (dimension 0: x 16 doubles)
};
struct TmaShared {
    unsigned long long full[kTmaBufs], done[kTmaBufs], empty[kTmaBufs];
    // which tile of this CTA the buffer holds (written by the loader before it arms `full`).  A parity wait cannot tell
    // "phase n + 1 is complete" from "phase n is not complete yet"; tiles k and k + 3 share a buffer but belong to different
    // teams, so a team that gets to tile k + 3 while the load of tile k is still in flight would fall through its wait.
    volatile unsigned long long seq[kTmaBufs];
    TileArgs A;
    // the bits a tile does NOT span, as runs of consecutive positions: tile index -> index of the tile's element 0 in a few
    // shifts (a bit-by-bit deposit over L positions, per thread and tile, was 10 % of the phase pass's stall samples)
    int n_runs;
    unsigned char run_in[12], run_w[12], run_out[12];
};
__device__ __forceinline__ void deposit_runs(TmaShared& sh, const int L) {       // one thread, once per CTA
    int n = 0, in = 0;
    for (int p = 0; p < L;) {
        if ((sh.A.mask >> p) & 1ull) { ++p; continue; }
        int w = 0;
        while (p + w < L && !((sh.A.mask >> (p + w)) & 1ull)) ++w;
        sh.run_in[n] = (unsigned char)in;
        sh.run_w[n] = (unsigned char)w;
        sh.run_out[n] = (unsigned char)p;
        ++n;
        in += w;
        p += w;
    }
    sh.n_runs = n;
}
__device__ __forceinline__ unsigned long long deposit(const TmaShared& sh, const unsigned long long t) {
    unsigned long long base = 0;
    for (int r = 0; r < sh.n_runs; ++r) base |= ((t >> sh.run_in[r]) & ((1ull << sh.run_w[r]) - 1ull)) << sh.run_out[r];
    return base;
}

__device__ __forceinline__ unsigned s_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void s_mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = s_u32(bar);
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_coords(const TmaGeom& g, unsigned long long base, int (&c)[5]) {
    // dimension 0: 128-byte rows (16 doubles); its coordinate carries the tile-index bits between bit 3 and the first run
    c[0] = (int)(((base >> g.start[0]) & ((1ull << g.nbits[0]) - 1ull)) << 4);
#pragma unroll
    for (int d = 1; d < 5; ++d) c[d] = d < g.rank ? (int)((base >> g.start[d]) & ((1ull << g.nbits[d]) - 1ull)) : 0;
}

template <bool SCALED, bool PHASE>
__global__ void __launch_bounds__(kTmaThreads, 1) k_slice_rx_tma(const __grid_constant__ CUtensorMap map, int L,
                                                                 const TileArgs* __restrict__ ta, const TmaGeom geom,
                                                                 const PhaseArgs* __restrict__ pa, unsigned long long high) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    unsigned char* smem_raw = smem_dyn + ((1024u - (s_u32(smem_dyn) & 1023u)) & 1023u);      // swizzled TMA boxes: 1 KiB aligned
    double2* tiles = reinterpret_cast<double2*>(smem_raw);
    __shared__ __align__(16) TmaShared sh;
    __shared__ PhaseTabs PT[1];                           // PHASE (contiguous tiles only): see k_slice_rx_tile
    __shared__ PhaseTile PQ[kTmaTeams][2];
    constexpr unsigned kTileBytes = (unsigned)(sizeof(double2) << kTileBits12);
    if (threadIdx.x == 0) {
        sh.A = *ta;
        deposit_runs(sh, L);
        if (PHASE) PT[0].has_ww = 0;
        for (int b = 0; b < kTmaBufs; ++b) {
            sh.seq[b] = ~0ull;
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(&sh.full[b])), "r"(1) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(&sh.done[b])), "r"(kThreads) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(&sh.empty[b])), "r"(1) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (PHASE) {
        phase_setup(PT[0], pa, sh.A, threadIdx.x);
        __syncthreads();
    }
    const unsigned long long n_tiles = 1ull << (L - kTileBits12);
    const unsigned long long mine = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   // tiles of this CTA
    const int wg = threadIdx.x / kThreads;
    if (wg < kTmaTeams) {
        // ------------------------------ consumer team ----------------------------------------------------------
        const int tid = threadIdx.x - wg * kThreads;
        const double pre = SCALED ? sh.A.post : 1.0;
        PhaseThread PH;
        int par = 0;                                                  // which PhaseTile of the team the current tile uses
        if (PHASE) {
            phase_thread_init(PT[0], PH, tid);
            if ((unsigned long long)wg < mine) {
                if (tid < 64) phase_tile_setup(PT[0], PQ[wg][0], (high << L) | deposit(sh, blockIdx.x + wg * gridDim.x), tid);
                team_bar(1 + wg);
            }
        }
        for (unsigned long long k = (unsigned long long)wg; k < mine; k += kTmaTeams) {
            const int b = (int)(k % kTmaBufs);
            double2* tile = tiles + ((size_t)b << kTileBits12);
            do {
                s_mbar_wait(&sh.full[b], (unsigned)((k / kTmaBufs) & 1ull));
            } while (sh.seq[b] != k);
            if (PHASE) {
                // [rotations left over from the previous step] [phase] [this step's rotations]; the cosine product of both
                // rotation sets rides on the phase
                if (sh.A.n_pre) tile_rounds<SCALED>(tile, kTileBits12, sh.A.n_pre, sh.A.pre_active, sh.A.pre_c, sh.A.pre_s, tid, 1 + wg, 1.0);
                phase_apply(PT[0], PQ[wg][par], PH, tile, tid, pre);
                team_bar(1 + wg);
                // the per-tile constants of the team's NEXT tile (its index deposited into the bits the tile does not span),
                // published by the barriers of the rotation rounds below
                if (k + kTmaTeams < mine && tid < 64)
                    phase_tile_setup(PT[0], PQ[wg][par ^ 1], (high << L) | deposit(sh, blockIdx.x + (k + kTmaTeams) * gridDim.x), tid);
                if (sh.A.n_active == 0) team_bar(1 + wg);
                par ^= 1;
            }
            tile_rounds<SCALED>(tile, kTileBits12, sh.A.n_active, sh.A.active, sh.A.c, sh.A.s, tid, 1 + wg, PHASE ? 1.0 : pre);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // this thread's tile writes -> the bulk store
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(&sh.done[b])) : "memory");
        }
    } else if (threadIdx.x == kTmaTeams * kThreads) {
        // ------------------------------ loader -------------------------------------------------------------------
        for (unsigned long long k = 0; k < mine; ++k) {
            const int b = (int)(k % kTmaBufs);
            const unsigned long long base = deposit(sh, blockIdx.x + k * gridDim.x);
            int c[5];
            tma_coords(geom, base, c);
            if (k >= kTmaBufs) s_mbar_wait(&sh.empty[b], (unsigned)(((k / kTmaBufs) - 1ull) & 1ull));   // the store of tile k - 3 has read it
            sh.seq[b] = k;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(&sh.full[b])), "r"(kTileBytes) : "memory");
            const unsigned dst = s_u32(tiles + ((size_t)b << kTileBits12)), bar = s_u32(&sh.full[b]);
            switch (geom.rank) {
                case 2: asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                     ::"r"(dst), "l"(&map), "r"(bar), "r"(c[0]), "r"(c[1]) : "memory"); break;
                case 3: asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                                     ::"r"(dst), "l"(&map), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]) : "memory"); break;
                case 4: asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                                     ::"r"(dst), "l"(&map), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]) : "memory"); break;
                default: asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                                      ::"r"(dst), "l"(&map), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory"); break;
            }
        }
    } else if (threadIdx.x == kTmaTeams * kThreads + 32) {
        // ------------------------------ storer -------------------------------------------------------------------
        for (unsigned long long k = 0; k < mine; ++k) {
            const int b = (int)(k % kTmaBufs);
            const unsigned long long base = deposit(sh, blockIdx.x + k * gridDim.x);
            int c[5];
            tma_coords(geom, base, c);
            s_mbar_wait(&sh.done[b], (unsigned)((k / kTmaBufs) & 1ull));
            const unsigned src = s_u32(tiles + ((size_t)b << kTileBits12));
            switch (geom.rank) {
                case 2: asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                                     ::"l"(&map), "r"(src), "r"(c[0]), "r"(c[1]) : "memory"); break;
                case 3: asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                                     ::"l"(&map), "r"(src), "r"(c[0]), "r"(c[1]), "r"(c[2]) : "memory"); break;
                case 4: asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                     ::"l"(&map), "r"(src), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]) : "memory"); break;
                default: asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                                      ::"l"(&map), "r"(src), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory"); break;
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the store has read the buffer: the loader may refill it
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(&sh.empty[b])) : "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");              // every store is complete before the kernel ends
    }
}

__global__ void __launch_bounds__(kThreads) k_slice_fill(double2* __restrict__ psi, size_t N, double amp) {
    for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < N; x += (size_t)gridDim.x * blockDim.x)
        psi[x] = make_double2(amp, 0.0);
}

// partial <psi|M|psi>, M = c0 + sum_e w_e z_a z_b evaluated on the fly (a 2^n table would not fit at n = 32)
__global__ void __launch_bounds__(kThreads) k_slice_energy(const double2* __restrict__ psi, int L, unsigned long long high,
                                                           const PhaseArgs* __restrict__ pa, double* __restrict__ partial) {
    __shared__ double w[kMaxPairs];
    __shared__ unsigned char ba[kMaxPairs], bb[kMaxPairs];
    __shared__ double red[kThreads / 32];
    const int n_zz = pa->n_zz;
    for (int e = threadIdx.x; e < n_zz; e += blockDim.x) { w[e] = pa->ang[e]; ba[e] = pa->a[e]; bb[e] = pa->b[e]; }
    __syncthreads();
    const double c0 = pa->c0;
    const size_t N = (size_t)1 << L;
    const unsigned long long hi = high << L;
    double acc = 0.0;
    for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < N; x += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long g = hi | x;
        double m = c0;
        for (int e = 0; e < n_zz; ++e) m += (((g >> ba[e]) ^ (g >> bb[e])) & 1ull) ? -w[e] : w[e];
        const double2 v = psi[x];
        acc = fma(m, v.x * v.x + v.y * v.y, acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < kThreads / 32 ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) partial[blockIdx.x] = v;
    }
}

// <psi|M|psi> partial with the same Gray-code walk: m(g ^ 1<<p) = m(g) - 2 sum_{pairs e containing p} w_e sigma_e(g).
// GrayArgs::c2 holds 2 w_e here.
__global__ void __launch_bounds__(kThreads) k_slice_energy_gray(const double2* __restrict__ psi, int L, unsigned long long high,
                                                                const PhaseArgs* __restrict__ pa, const GrayArgs* __restrict__ ga,
                                                                double* __restrict__ partial) {
    __shared__ double w[kMaxPairs];
    __shared__ unsigned char ba[kMaxPairs], bb[kMaxPairs];
    __shared__ double red[kThreads / 32];
    __shared__ GrayArgs G;
    const int n_zz = pa->n_zz;
    for (int e = threadIdx.x; e < n_zz; e += blockDim.x) { w[e] = pa->ang[e]; ba[e] = pa->a[e]; bb[e] = pa->b[e]; }
    if (threadIdx.x == 0) G = *ga;
    __syncthreads();
    const double c0 = pa->c0;
    const int B = G.B, S = G.S;
    const size_t n_low = (size_t)1 << S;
    const unsigned long long hi = high << L;
    double acc = 0.0;
    for (size_t low = blockIdx.x * (size_t)blockDim.x + threadIdx.x; low < n_low; low += (size_t)gridDim.x * blockDim.x) {
        unsigned long long g = hi | low;
        double m = c0;
        for (int e = 0; e < n_zz; ++e) m += (((g >> ba[e]) ^ (g >> bb[e])) & 1ull) ? -w[e] : w[e];
        {
            const double2 v = psi[low];
            acc = fma(m, v.x * v.x + v.y * v.y, acc);
        }
        for (int k = 1; k < (1 << B); ++k) {
            const int j = __ffs(k) - 1, p = S + j;
            const int d = G.deg[j];
            for (int q = 0; q < d; ++q) {
                const bool differ = (((g >> G.other[j][q]) ^ (g >> p)) & 1ull) != 0;
                m += differ ? G.c2[j][q] : -G.c2[j][q];                      // sigma flips sign: -w -> +w adds 2w
            }
            g ^= 1ull << p;
            const double2 v = psi[(size_t)(g & ((1ull << L) - 1ull))];
            acc = fma(m, v.x * v.x + v.y * v.y, acc);
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < kThreads / 32 ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) partial[blockIdx.x] = v;
    }
}

int grid_for(const dq_context* ctx, size_t work) {
    const size_t b = (work + kThreads - 1) / kThreads, cap = (size_t)ctx->prop.multiProcessorCount * 16;
    return (int)std::max<size_t>(1, std::min(b, cap));
}

int fill_args(PhaseArgs& h, int n_total, int n_zz, const int32_t* pair_bits, const double* vals, double c0) {
    DQ_REQUIRE(n_zz >= 0 && n_zz <= kMaxPairs, "slice: n_zz=%d outside [0,%d]", n_zz, kMaxPairs);
    DQ_REQUIRE(n_zz == 0 || (pair_bits && vals), "slice: NULL pair table");
    h.n_zz = n_zz;
    h.c0 = c0;
    for (int e = 0; e < n_zz; ++e) {
        const int a = pair_bits[2 * e], b = pair_bits[2 * e + 1];
        DQ_REQUIRE(a >= 0 && b >= 0 && a < n_total && b < n_total && a != b, "slice: pair %d = bits (%d,%d) invalid for %d qubits", e, a, b, n_total);
        DQ_REQUIRE(isfinite(vals[e]), "slice: non-finite value for pair %d", e);
        h.a[e] = (unsigned char)a;
        h.b[e] = (unsigned char)b;
        h.ang[e] = vals[e];
    }
    DQ_REQUIRE(isfinite(c0), "slice: non-finite constant");
    return DQ_OK;
}

int upload_args(dq_context* ctx, const PhaseArgs& h, PhaseArgs** d_out) {
    // a small ring so that consecutive asynchronous launches never share a table
    if (!ctx->slice_ring) DQ_CUDA(cudaMalloc(&ctx->slice_ring, sizeof(PhaseArgs) * kRingSlots));
    PhaseArgs* d = reinterpret_cast<PhaseArgs*>(ctx->slice_ring) + (ctx->slice_cursor++ % kRingSlots);
    if ((ctx->slice_cursor % kRingSlots) == 0) DQ_CUDA(cudaStreamSynchronize(ctx->stream));      // ring wrapped: drain
    DQ_CUDA(cudaMemcpyAsync(d, &h, sizeof(PhaseArgs), cudaMemcpyHostToDevice, ctx->stream));
    *d_out = d;
    return DQ_OK;
}

// pair lists of the top B local bits; false when a bit has more than kGrayDeg pairs (dense graphs: use the plain kernels)
bool build_gray(const PhaseArgs& h, int L, bool energy, GrayArgs& gh) {
    memset(&gh, 0, sizeof(gh));
    gh.B = std::min(kGrayBits, L);
    gh.S = L - gh.B;
    if (L < 10) return false;                              // small slices: the plain kernel is a single wave anyway
    for (int j = 0; j < gh.B; ++j)
        for (int e = 0; e < h.n_zz; ++e) {
            const int p = gh.S + j;
            if (h.a[e] != p && h.b[e] != p) continue;
            if (gh.deg[j] == kGrayDeg) return false;
            gh.other[j][gh.deg[j]] = h.a[e] == p ? h.b[e] : h.a[e];
            gh.c2[j][gh.deg[j]] = energy ? 2.0 * h.ang[e] : cos(2.0 * h.ang[e]);
            gh.s2[j][gh.deg[j]] = energy ? 0.0 : sin(2.0 * h.ang[e]);
            ++gh.deg[j];
        }
    return true;
}

int upload_gray(dq_context* ctx, const GrayArgs& gh, GrayArgs** d_out) {
    static_assert(sizeof(GrayArgs) <= sizeof(PhaseArgs), "ring slot too small");
    GrayArgs* dg = reinterpret_cast<GrayArgs*>(reinterpret_cast<PhaseArgs*>(ctx->slice_ring) + (ctx->slice_cursor++ % kRingSlots));
    if ((ctx->slice_cursor % kRingSlots) == 0) DQ_CUDA(cudaStreamSynchronize(ctx->stream));
    DQ_CUDA(cudaMemcpyAsync(dg, &gh, sizeof(GrayArgs), cudaMemcpyHostToDevice, ctx->stream));
    *d_out = dg;
    return DQ_OK;
}

}  // namespace

extern "C" {

int dq_slice_fill_uniform(dq_context* ctx, void* psi_dev, int L, int n_total) {
    DQ_REQUIRE(ctx && psi_dev, "NULL argument");
    DQ_REQUIRE(L >= 1 && L <= 33 && n_total >= L && n_total <= 40, "dq_slice_fill_uniform: L=%d n=%d", L, n_total);
    DQ_TRY(ctx->set_device());
    const size_t N = (size_t)1 << L;
    k_slice_fill<<<grid_for(ctx, N), kThreads, 0, ctx->stream>>>((double2*)psi_dev, N, exp2(-0.5 * n_total));
    ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

int dq_slice_phase(dq_context* ctx, void* psi_dev, int L, uint64_t high_bits, int n_total, int n_zz,
                   const int32_t* pair_bits, const double* angles) {
    DQ_REQUIRE(ctx && psi_dev && angles, "NULL argument");
    DQ_REQUIRE(L >= 1 && L <= 33 && n_total >= L && n_total <= 40, "dq_slice_phase: L=%d n=%d", L, n_total);
    DQ_REQUIRE((high_bits >> (n_total - L)) == 0, "dq_slice_phase: high_bits do not fit %d global bits", n_total - L);
    DQ_TRY(ctx->set_device());
    PhaseArgs h;
    DQ_TRY(fill_args(h, n_total, n_zz, pair_bits, angles + 1, angles[0]));
    PhaseArgs* d;
    DQ_TRY(upload_args(ctx, h, &d));
    // Gray-code walk over the top B local bits when every one of them has a short pair list (any sparse graph)
    GrayArgs gh;
    const bool gray_ok = build_gray(h, L, false, gh);
    if (gray_ok) {
        GrayArgs* dg;
        DQ_TRY(upload_gray(ctx, gh, &dg));
        k_slice_phase_gray<<<grid_for(ctx, (size_t)1 << gh.S), kThreads, 0, ctx->stream>>>((double2*)psi_dev, L, high_bits, d, dg);
    } else {
        k_slice_phase<<<grid_for(ctx, (size_t)1 << L), kThreads, 0, ctx->stream>>>((double2*)psi_dev, L, high_bits, d);
    }
    ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

int dq_slice_rx(dq_context* ctx, void* psi_dev, int L, int bit, double theta) {
    DQ_REQUIRE(ctx && psi_dev, "NULL argument");
    DQ_REQUIRE(L >= 1 && L <= 33 && bit >= 0 && bit < L, "dq_slice_rx: bit %d is not local (L=%d)", bit, L);
    DQ_REQUIRE(isfinite(theta), "dq_slice_rx: non-finite angle");
    DQ_TRY(ctx->set_device());
    k_slice_rx<<<grid_for(ctx, (size_t)1 << (L - 1)), kThreads, 0, ctx->stream>>>((double2*)psi_dev, L, bit, cos(theta), sin(theta));
    ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

}  // extern "C"

namespace {
// All rotations of a step; d_phase != NULL: the step's diagonal phase is applied by the first pass if that pass is a
// contiguous 12-bit tile pass (*phase_done says whether it was).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn slice_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
        cudaGetLastError();
    }
    return fn;
}

struct Scatter { int g, rank; void* const* peer; };
struct Rot { int bit; double theta; };

// dq_slice_plan: the planner runs for real, the launches are written down instead of made (no device needed)
struct PlanRec {
    std::vector<int32_t> rows;      // per launch: step, T, lo, n_pre, phase, n_rot, scatter, mask low 32, mask high 32
    bool assume_tma;
    int step;
};
thread_local PlanRec* g_rec = nullptr;

// One tile geometry and the rotation targets it carries.  A tile is 2^T amplitudes spanned by T physical bits pos[0] < pos[1] < ...:
// the `lo` lowest are bits 0 .. lo-1 (contiguous 16 * 2^lo byte runs in global memory), the others are targets.
struct TileSet {
    int T, lo;
    unsigned char pos[12];
    unsigned long long mask;
    int n_targets;
    int target[12];                 // physical bits, ascending; every one is a pos[]
    bool contiguous() const { return lo == T; }
    bool has(int bit) const {
        for (int i = 0; i < n_targets; ++i)
            if (target[i] == bit) return true;
        return false;
    }
};

int tile_bits_max() {               // 2^12 amplitudes (64 KiB) unless DQ_SLICE_TILE_BITS says otherwise (8..12; experiments)
    static int tile_bits = 0;
    if (!tile_bits) {
        const char* env = getenv("DQ_SLICE_TILE_BITS");
        const int v = env ? atoi(env) : 12;
        tile_bits = (v >= 8 && v <= 12) ? v : 12;
    }
    return tile_bits;
}

// The tiles that cover the (ascending) target bits.  Every target below bit Tmax goes into ONE contiguous tile (bits 0 .. Tmax-1).
// High targets: the tile is 2^Tmax amplitudes - `lo` contiguous low bits + the targets of the pass; the targets are split evenly
// over the fewest passes that leave lo >= 3 (128-byte runs: one row of the TMA tile), so a pass with few targets gets long
// contiguous runs instead of a small tile.
std::vector<TileSet> plan_sets(int L, const std::vector<int>& bits) {
    std::vector<TileSet> sets;
    const int Tmax = std::min(tile_bits_max(), L);
    static int min_lo = 0;
    if (!min_lo) {
        const char* env = getenv("DQ_SLICE_MIN_LO");
        const int v = env ? atoi(env) : 3;
        min_lo = (v >= 1 && v <= 8) ? v : 3;
    }
    size_t next = 0;
    if (next < bits.size() && bits[next] < Tmax) {
        TileSet S;
        memset(&S, 0, sizeof(S));
        S.T = S.lo = Tmax;
        for (int i = 0; i < Tmax; ++i) S.pos[i] = (unsigned char)i;
        while (next < bits.size() && bits[next] < Tmax) S.target[S.n_targets++] = bits[next++];
        sets.push_back(S);
    }
    const int rest = (int)(bits.size() - next);
    if (rest > 0) {
        const int cap = std::max(1, Tmax - std::min(min_lo, Tmax - 1));
        const int passes = (rest + cap - 1) / cap;
        for (int k = 0; k < passes; ++k) {
            const int cnt = rest / passes + (k < rest % passes ? 1 : 0);
            TileSet S;
            memset(&S, 0, sizeof(S));
            S.lo = Tmax - cnt;
            for (int i = 0; i < S.lo; ++i) S.pos[i] = (unsigned char)i;
            S.T = S.lo;
            for (int i = 0; i < cnt; ++i) {
                S.pos[S.T++] = (unsigned char)bits[next];
                S.target[S.n_targets++] = bits[next++];
            }
            sets.push_back(S);
        }
    }
    for (auto& S : sets)
        for (int i = 0; i < S.T; ++i) S.mask |= 1ull << S.pos[i];
    return sets;
}

// Tensor view of a tile set over the slice (see k_slice_rx_tma); false: not expressible with rank <= 5 (or no driver entry).
bool tile_geom(const TileSet& S, int L, TmaGeom* g, cuuint64_t* dims, cuuint64_t* strides, cuuint32_t* box) {
    if (S.T != kTileBits12 || L < kTileBits12 || S.lo < 3) return false;
    // runs of consecutive tile bits above bit 2 (bits 0..2 are dimension 0), at most 8 bits each (box extent <= 256)
    int run_start[8], run_len[8], n_runs = 0;
    for (int i = 3; i < S.T;) {
        int len = 1;
        while (i + len < S.T && S.pos[i + len] == S.pos[i] + len && len < 8) ++len;
        if (n_runs == 4) return false;
        run_start[n_runs] = S.pos[i];
        run_len[n_runs] = len;
        ++n_runs;
        i += len;
    }
    if (n_runs == 0) return false;
    // dimension 0: rows of 8 amplitudes = 16 doubles = 128 bytes (tile bits 0..2), spanning up to the first run: with lo = 3 the
    // first run starts above bit 3 and the tile-index bits in between select the row
    dims[0] = (cuuint64_t)16 << (run_start[0] - 3);
    box[0] = 16;
    g->rank = 1 + n_runs;
    g->start[0] = 3;
    g->nbits[0] = run_start[0] - 3;
    for (int r = 0; r < n_runs; ++r) {
        const int end = r + 1 < n_runs ? run_start[r + 1] : L;          // the dimension spans up to the next run
        dims[1 + r] = (cuuint64_t)1 << (end - run_start[r]);
        strides[r] = (cuuint64_t)16 << run_start[r];
        box[1 + r] = (cuuint32_t)1 << run_len[r];
        g->start[1 + r] = run_start[r];
        g->nbits[1 + r] = end - run_start[r];
    }
    for (int d = g->rank; d < 5; ++d) { g->start[d] = 0; g->nbits[d] = 0; }
    return true;
}

bool tma_usable(const TileSet& S, int L) {
    TmaGeom g;
    cuuint64_t dims[5], strides[4];
    cuuint32_t box[5];
    if (g_rec) return g_rec->assume_tma && tile_geom(S, L, &g, dims, strides, box);
    return !getenv("DQ_SLICE_NO_TMA") && slice_encode_fn() && tile_geom(S, L, &g, dims, strides, box);
}

bool make_tile_map(const TileSet& S, int L, void* psi, CUtensorMap* map, TmaGeom* g) {
    EncodeTiledFn enc = slice_encode_fn();
    cuuint64_t dims[5], strides[4];
    cuuint32_t box[5], ones[5] = {1, 1, 1, 1, 1};
    if (!enc || !tile_geom(S, L, g, dims, strides, box)) return false;
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)g->rank, psi, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Do the kernels' phase tables hold this pair list for this tile?  (12-bit tiles only; every tile bit at most kGrayDeg pairs,
// at most kMaxIn pairs inside the tile.)
bool phase_fits(const PhaseArgs& h, const TileSet& S) {
    if (S.T != kTileBits) return false;
    int n_in = 0, deg[kTileBits] = {0};
    for (int e = 0; e < h.n_zz; ++e) {
        int ta = -1, tb = -1;
        for (int i = 0; i < S.T; ++i) {
            if (S.pos[i] == h.a[e]) ta = i;
            if (S.pos[i] == h.b[e]) tb = i;
        }
        if (ta >= 0) ++deg[ta];
        if (tb >= 0) ++deg[tb];
        n_in += (ta >= 0 && tb >= 0) ? 1 : 0;
    }
    for (int k = 0; k < kTileBits; ++k)
        if (deg[k] > kGrayDeg) return false;
    return n_in <= kMaxIn;
}

// Can launch_pass carry the phase (and, with `pre`, the previous step's rotations) on this set?  A scatter pass runs on the
// cp.async kernel, which knows the phase for the contiguous tile only and no pre-rotations.
bool can_carry_phase(const PhaseArgs& h, const TileSet& S, int L, bool pre, bool scatter) {
    if (!phase_fits(h, S)) return false;
    if (!scatter && tma_usable(S, L)) return true;
    return !pre && S.contiguous() && !getenv("DQ_SLICE_NO_PHASE_FUSION");
}

// ONE pass over the slice: [pre: rotations left over from the previous step] [the diagonal phase, d_phase != NULL] [rot], all
// on the bits of tile set S (rot / pre ascending in their bit, every bit a target of S); sc: the stores of this pass are the
// exchange.  Callers make sure the combination is supported (can_carry_phase).
int launch_pass(dq_context* ctx, void* psi_dev, int L, const TileSet& S, const std::vector<Rot>& rot, const std::vector<Rot>& pre,
                const PhaseArgs* d_phase, unsigned long long high_bits, const Scatter* sc) {
    if (g_rec) {
        const int32_t row[9] = {g_rec->step, S.T, S.lo, (int32_t)pre.size(), d_phase ? 1 : 0, (int32_t)rot.size(), sc ? 1 : 0,
                                (int32_t)(S.mask & 0xffffffffull), (int32_t)(S.mask >> 32)};
        g_rec->rows.insert(g_rec->rows.end(), row, row + 9);
        return DQ_OK;
    }
    TileArgs h;
    memset(&h, 0, sizeof(h));
    h.T = S.T;
    h.lo = S.lo;
    h.mask = S.mask;
    memcpy(h.pos, S.pos, sizeof(h.pos));
    auto tile_bit = [&](int bit) {
        for (int i = 0; i < S.T; ++i)
            if (S.pos[i] == bit) return i;
        return -1;
    };
    for (const Rot& r : rot) {
        const int tb = tile_bit(r.bit);
        DQ_REQUIRE(tb >= 0 && h.n_active < 12, "slice pass: internal error: bit %d is not in the tile", r.bit);
        h.active[h.n_active] = (unsigned char)tb;
        h.c[h.n_active] = cos(r.theta);
        h.s[h.n_active] = sin(r.theta);
        ++h.n_active;
    }
    for (const Rot& r : pre) {
        const int tb = tile_bit(r.bit);
        DQ_REQUIRE(tb >= 0 && h.n_pre < 12, "slice pass: internal error: bit %d is not in the tile", r.bit);
        h.pre_active[h.n_pre] = (unsigned char)tb;
        h.pre_c[h.n_pre] = cos(r.theta);
        h.pre_s[h.n_pre] = sin(r.theta);
        ++h.n_pre;
    }
    if (sc) {
        h.scatter_g = sc->g;
        h.scatter_rank = sc->rank;
        for (int j = 0; j < (1 << sc->g); ++j) h.peer[j] = (double2*)sc->peer[j];
    }
    // scaled form unless a rotation angle sits close to pi/2 (|tan| large: the cosine product would lose digits)
    h.scaled = 1;
    h.post = 1.0;
    for (int k = 0; k < h.n_active; ++k)
        if (fabs(h.c[k]) < 0.05) h.scaled = 0;
    for (int k = 0; k < h.n_pre; ++k)
        if (fabs(h.pre_c[k]) < 0.05) h.scaled = 0;
    if (h.scaled) {
        for (int k = 0; k < h.n_active; ++k) {
            h.post *= h.c[k];
            h.s[k] = h.s[k] / h.c[k];
            h.c[k] = 1.0;
        }
        for (int k = 0; k < h.n_pre; ++k) {
            h.post *= h.pre_c[k];
            h.pre_s[k] = h.pre_s[k] / h.pre_c[k];
            h.pre_c[k] = 1.0;
        }
    }
    if (!ctx->slice_ring) DQ_CUDA(cudaMalloc(&ctx->slice_ring, sizeof(PhaseArgs) * kRingSlots));
    static_assert(sizeof(TileArgs) <= sizeof(PhaseArgs), "ring slot too small");
    TileArgs* d = reinterpret_cast<TileArgs*>(reinterpret_cast<PhaseArgs*>(ctx->slice_ring) + (ctx->slice_cursor++ % kRingSlots));
    if ((ctx->slice_cursor % kRingSlots) == 0) DQ_CUDA(cudaStreamSynchronize(ctx->stream));
    DQ_CUDA(cudaMemcpyAsync(d, &h, sizeof(TileArgs), cudaMemcpyHostToDevice, ctx->stream));
    const unsigned long long n_tiles = 1ull << (L - h.T);
    double2* psi = (double2*)psi_dev;
    if (!sc && !getenv("DQ_SLICE_NO_TMA")) {
        // TMA tiles, three buffers per SM (k_slice_rx_tma); the map is a kernel parameter, nothing to keep alive
        CUtensorMap map;
        TmaGeom geom;
        if (make_tile_map(S, L, psi_dev, &map, &geom)) {
            const size_t tsmem = (size_t)kTmaBufs * (sizeof(double2) << kTileBits12) + 1024;
            if (!ctx->slice_tma_attr) {         // per device (a context is one device)
                DQ_CUDA(cudaFuncSetAttribute(k_slice_rx_tma<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
                DQ_CUDA(cudaFuncSetAttribute(k_slice_rx_tma<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
                DQ_CUDA(cudaFuncSetAttribute(k_slice_rx_tma<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
                DQ_CUDA(cudaFuncSetAttribute(k_slice_rx_tma<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
                ctx->slice_tma_attr = true;
            }
            const unsigned tgrid = (unsigned)std::min<unsigned long long>(n_tiles, (unsigned long long)ctx->prop.multiProcessorCount);
            if (d_phase) {
                if (h.scaled) k_slice_rx_tma<true, true><<<tgrid, kTmaThreads, tsmem, ctx->stream>>>(map, L, d, geom, d_phase, high_bits);
                else k_slice_rx_tma<false, true><<<tgrid, kTmaThreads, tsmem, ctx->stream>>>(map, L, d, geom, d_phase, high_bits);
            }
            else if (h.scaled) k_slice_rx_tma<true, false><<<tgrid, kTmaThreads, tsmem, ctx->stream>>>(map, L, d, geom, nullptr, 0);
            else k_slice_rx_tma<false, false><<<tgrid, kTmaThreads, tsmem, ctx->stream>>>(map, L, d, geom, nullptr, 0);
            ctx->launches++;
            DQ_CUDA(cudaGetLastError());
            return DQ_OK;
        }
    }
    DQ_REQUIRE(pre.empty() && (!d_phase || (S.contiguous() && S.T == kTileBits)),
               "slice pass: internal error: this pass needs the TMA tile kernel (T=%d lo=%d)", S.T, S.lo);
    {   // the attribute is per device and cheap to set: no process-wide "done" flag (two contexts, two devices)
        const int max_smem = (int)(sizeof(double2) << 12);
        DQ_CUDA(cudaFuncSetAttribute(k_slice_rx_tile<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        DQ_CUDA(cudaFuncSetAttribute(k_slice_rx_tile<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        DQ_CUDA(cudaFuncSetAttribute(k_slice_rx_tile<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        DQ_CUDA(cudaFuncSetAttribute(k_slice_rx_tile<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        DQ_CUDA(cudaFuncSetAttribute(k_slice_rx_tile<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        DQ_CUDA(cudaFuncSetAttribute(k_slice_rx_tile<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    }
    const size_t smem = sizeof(double2) << h.T;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, ((size_t)200 << 10) / std::max<size_t>(smem, 1)));
    const unsigned grid = (unsigned)std::min<unsigned long long>(n_tiles, (unsigned long long)ctx->prop.multiProcessorCount * per_sm);
    if (d_phase) {
        if (h.scaled) k_slice_rx_tile<true, true, true><<<grid, kThreads, smem, ctx->stream>>>(psi, L, d, d_phase, high_bits);
        else k_slice_rx_tile<true, false, true><<<grid, kThreads, smem, ctx->stream>>>(psi, L, d, d_phase, high_bits);
    }
    else if (S.contiguous() && h.scaled) k_slice_rx_tile<true, true, false><<<grid, kThreads, smem, ctx->stream>>>(psi, L, d, nullptr, 0);
    else if (S.contiguous()) k_slice_rx_tile<true, false, false><<<grid, kThreads, smem, ctx->stream>>>(psi, L, d, nullptr, 0);
    else if (h.scaled) k_slice_rx_tile<false, true, false><<<grid, kThreads, smem, ctx->stream>>>(psi, L, d, nullptr, 0);
    else k_slice_rx_tile<false, false, false><<<grid, kThreads, smem, ctx->stream>>>(psi, L, d, nullptr, 0);
    ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

// (bit, theta) lists: validated, ascending in the bit
int sorted_rots(const char* what, int L, int count, const int32_t* bits, const double* thetas, std::vector<Rot>& out) {
    DQ_REQUIRE(count == 0 || (bits && thetas), "%s: NULL argument", what);
    DQ_REQUIRE(count >= 0 && count <= L, "%s: L=%d count=%d", what, L, count);
    out.resize(count);
    unsigned long long seen = 0;
    for (int i = 0; i < count; ++i) {
        DQ_REQUIRE(bits[i] >= 0 && bits[i] < L, "%s: bit %d is not local (L=%d)", what, bits[i], L);
        DQ_REQUIRE(!((seen >> bits[i]) & 1ull), "%s: bit %d listed twice", what, bits[i]);
        DQ_REQUIRE(isfinite(thetas[i]), "%s: non-finite angle", what);
        seen |= 1ull << bits[i];
        out[i] = Rot{bits[i], thetas[i]};
    }
    std::sort(out.begin(), out.end(), [](const Rot& a, const Rot& b) { return a.bit < b.bit; });
    return DQ_OK;
}

std::vector<Rot> rots_of(const TileSet& S, const std::vector<Rot>& all) {
    std::vector<Rot> r;
    for (const Rot& x : all)
        if (S.has(x.bit)) r.push_back(x);
    return r;
}

// All rotations of a step, every tile set one pass; sc: the last pass carries the exchange.
int rx_many_impl(dq_context* ctx, void* psi_dev, int L, int count, const int32_t* bits, const double* thetas, const Scatter* sc = nullptr) {
    DQ_REQUIRE(g_rec || (ctx && psi_dev), "NULL argument");
    DQ_REQUIRE(L >= 1 && L <= 33, "dq_slice_rx_many: L=%d", L);
    if (!g_rec) DQ_TRY(ctx->set_device());
    std::vector<Rot> rot;
    DQ_TRY(sorted_rots("dq_slice_rx_many", L, count, bits, thetas, rot));
    std::vector<int> tb(rot.size());
    for (size_t i = 0; i < rot.size(); ++i) tb[i] = rot[i].bit;
    const std::vector<TileSet> sets = plan_sets(L, tb);
    const std::vector<Rot> none;
    for (size_t k = 0; k < sets.size(); ++k)
        DQ_TRY(launch_pass(ctx, psi_dev, L, sets[k], rots_of(sets[k], rot), none, nullptr, 0, (sc && k + 1 == sets.size()) ? sc : nullptr));
    return DQ_OK;
}

// One product-formula step on a slice:  [pre: rotations still owed to the previous step] [phase, angles != NULL] [rot] (+ the
// exchange on the last pass).  `skip` >= 0: the rotations of tile set `skip` are NOT applied (the caller passes them as `pre`
// of the next step); `first` >= 0: the tile set that must carry pre + phase (the caller has checked can_carry_phase).
// Without `first` the function picks the set itself; when no set can carry the work the pieces run as separate passes.
int step_impl(dq_context* ctx, void* psi_dev, int L, unsigned long long high_bits, const PhaseArgs* h_phase, const std::vector<Rot>& pre,
              const std::vector<Rot>& rot, const std::vector<TileSet>& sets, int first, int skip, const Scatter* sc,
              int n_total, int n_zz, const int32_t* pair_bits, const double* angles) {
    const std::vector<Rot> none;
    const bool only_set_scatters = sc && sets.size() == 1;
    bool pre_fused = !pre.empty();
    if (first < 0 && h_phase && !pre.empty()) {                   // the set that holds every pre bit among its targets
        for (size_t k = 0; k < sets.size() && first < 0; ++k) {
            bool all = true;
            for (const Rot& r : pre) all = all && sets[k].has(r.bit);
            if (all && (int)k != skip && can_carry_phase(*h_phase, sets[k], L, true, only_set_scatters)) first = (int)k;
        }
    }
    if (first < 0 && !pre.empty()) {                              // no such set: the owed rotations as passes of their own
        std::vector<int32_t> pb(pre.size());
        std::vector<double> pt(pre.size());
        for (size_t i = 0; i < pre.size(); ++i) { pb[i] = pre[i].bit; pt[i] = pre[i].theta; }
        DQ_TRY(rx_many_impl(ctx, psi_dev, L, (int)pre.size(), pb.data(), pt.data()));
        pre_fused = false;
    }
    if (first < 0 && h_phase) {
        for (size_t k = 0; k < sets.size() && first < 0; ++k)
            if ((int)k != skip && can_carry_phase(*h_phase, sets[k], L, false, only_set_scatters)) first = (int)k;
    }
    if (first < 0 && h_phase) {
        if (g_rec) {                                              // the stand-alone phase pass
            const int32_t row[9] = {g_rec->step, 0, 0, 0, 1, 0, 0, 0, 0};
            g_rec->rows.insert(g_rec->rows.end(), row, row + 9);
        } else {
            DQ_TRY(dq_slice_phase(ctx, psi_dev, L, high_bits, n_total, n_zz, pair_bits, angles));
        }
    }
    PhaseArgs* d_phase = nullptr;
    if (first >= 0 && h_phase) {
        if (g_rec) d_phase = reinterpret_cast<PhaseArgs*>(sizeof(PhaseArgs));      // never dereferenced
        else DQ_TRY(upload_args(ctx, *h_phase, &d_phase));
    }
    // launch order: `first`, then the others; the last one launched carries the exchange
    std::vector<int> order;
    if (first >= 0) order.push_back(first);
    for (size_t k = 0; k < sets.size(); ++k)
        if ((int)k != first && (int)k != skip) order.push_back((int)k);
    if (sc && order.size() > 2 - (first < 0 ? 1 : 0)) {
        // the exchange rides on the pass with the longest contiguous runs (the contiguous tile: 64 KiB per remote write burst)
        const size_t from = first >= 0 ? 1 : 0;
        size_t best = from;
        for (size_t i = from; i < order.size(); ++i)
            if (sets[order[i]].lo > sets[order[best]].lo) best = i;
        std::swap(order[best], order.back());
    }
    for (size_t i = 0; i < order.size(); ++i) {
        const int k = order[i];
        const bool lead = k == first;
        DQ_TRY(launch_pass(ctx, psi_dev, L, sets[k], rots_of(sets[k], rot), (lead && pre_fused) ? pre : none, lead ? d_phase : nullptr, high_bits,
                           (sc && i + 1 == order.size()) ? sc : nullptr));
    }
    DQ_REQUIRE(!sc || !order.empty(), "slice step: needs at least one rotation to carry the exchange");
    return DQ_OK;
}
}  // namespace

extern "C" {

int dq_slice_rx_many(dq_context* ctx, void* psi_dev, int L, int count, const int32_t* bits, const double* thetas) {
    return rx_many_impl(ctx, psi_dev, L, count, bits, thetas);
}

int dq_slice_rx_many_scatter(dq_context* ctx, void* psi_dev, int L, int count, const int32_t* bits, const double* thetas,
                             int g, int rank, void* const* peer_recv) {
    DQ_REQUIRE(ctx && psi_dev && peer_recv, "NULL argument");
    DQ_REQUIRE(g >= 1 && (1 << g) <= kMaxPeers && g <= L && rank >= 0 && rank < (1 << g), "dq_slice_rx_many_scatter: g=%d rank=%d L=%d", g, rank, L);
    DQ_REQUIRE(count >= 1, "dq_slice_rx_many_scatter: needs at least one rotation to carry the exchange");
    for (int j = 0; j < (1 << g); ++j) DQ_REQUIRE(peer_recv[j] != nullptr, "dq_slice_rx_many_scatter: NULL receive buffer of rank %d", j);
    Scatter sc{g, rank, peer_recv};
    return rx_many_impl(ctx, psi_dev, L, count, bits, thetas, &sc);
}

int dq_ipc_export(dq_context* ctx, void* dev_ptr, void* handle64_out, uint64_t* offset_out) {
    DQ_REQUIRE(ctx && dev_ptr && handle64_out && offset_out, "NULL argument");
    DQ_TRY(ctx->set_device());
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CUdeviceptr base = 0;
    size_t size = 0;
    // the handle names the whole allocation the pointer lives in (a caching allocator may hand out an interior pointer)
    cudaPointerAttributes attr;
    DQ_CUDA(cudaPointerGetAttributes(&attr, dev_ptr));
    DQ_REQUIRE(attr.type == cudaMemoryTypeDevice, "dq_ipc_export: not a device pointer");
    {
        typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        DQ_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q));
        DQ_REQUIRE(q == cudaDriverEntryPointSuccess && f, "dq_ipc_export: cuMemGetAddressRange is not available");
        if (reinterpret_cast<RangeFn>(f)(&base, &size, (CUdeviceptr)dev_ptr) != CUDA_SUCCESS) {
            dq::set_error("dq_ipc_export: cuMemGetAddressRange failed");
            return DQ_ERR_CUDA;
        }
    }
    DQ_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64_out, (void*)base));
    *offset_out = (uint64_t)((CUdeviceptr)dev_ptr - base);
    return DQ_OK;
}

int dq_ipc_open(dq_context* ctx, const void* handle64, uint64_t offset, void** ptr_out) {
    DQ_REQUIRE(ctx && handle64 && ptr_out, "NULL argument");
    DQ_TRY(ctx->set_device());
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* base = nullptr;
    DQ_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr_out = (char*)base + offset;
    return DQ_OK;
}

int dq_ipc_close(dq_context* ctx, void* ptr, uint64_t offset) {
    DQ_REQUIRE(ctx && ptr, "NULL argument");
    DQ_TRY(ctx->set_device());
    DQ_CUDA(cudaIpcCloseMemHandle((char*)ptr - offset));
    return DQ_OK;
}

// pre-rotations (may be none), phase, rotations, exchange (may be none): see step_impl
static int slice_step(const char* what, dq_context* ctx, void* psi_dev, int L, uint64_t high_bits, int n_total, int n_zz,
                      const int32_t* pair_bits, const double* angles, int n_pre, const int32_t* pre_bits, const double* pre_thetas,
                      int count, const int32_t* bits, const double* thetas, const Scatter* sc) {
    DQ_REQUIRE((g_rec || (ctx && psi_dev)) && angles, "NULL argument");
    DQ_REQUIRE(L >= 1 && L <= 33 && n_total >= L && n_total <= 40, "%s: L=%d n=%d", what, L, n_total);
    DQ_REQUIRE((high_bits >> (n_total - L)) == 0, "%s: high_bits do not fit %d global bits", what, n_total - L);
    if (!g_rec) DQ_TRY(ctx->set_device());
    std::vector<Rot> pre, rot;
    DQ_TRY(sorted_rots(what, L, n_pre, pre_bits, pre_thetas, pre));
    DQ_TRY(sorted_rots(what, L, count, bits, thetas, rot));
    PhaseArgs h;
    DQ_TRY(fill_args(h, n_total, n_zz, pair_bits, angles + 1, angles[0]));
    std::vector<int> tb(rot.size());
    for (size_t i = 0; i < rot.size(); ++i) tb[i] = rot[i].bit;
    const std::vector<TileSet> sets = plan_sets(L, tb);
    return step_impl(ctx, psi_dev, L, high_bits, &h, pre, rot, sets, -1, -1, sc, n_total, n_zz, pair_bits, angles);
}

static int check_scatter(const char* what, int L, int count, int g, int rank, void* const* peer_recv) {
    DQ_REQUIRE(peer_recv, "NULL argument");
    DQ_REQUIRE(g >= 1 && (1 << g) <= kMaxPeers && g <= L && rank >= 0 && rank < (1 << g), "%s: g=%d rank=%d L=%d", what, g, rank, L);
    DQ_REQUIRE(count >= 1, "%s: needs at least one rotation to carry the exchange", what);
    for (int j = 0; j < (1 << g); ++j) DQ_REQUIRE(peer_recv[j] != nullptr, "%s: NULL receive buffer of rank %d", what, j);
    return DQ_OK;
}

int dq_slice_phase_rx_many(dq_context* ctx, void* psi_dev, int L, uint64_t high_bits, int n_total, int n_zz,
                           const int32_t* pair_bits, const double* angles, int count, const int32_t* bits, const double* thetas) {
    return slice_step("dq_slice_phase_rx_many", ctx, psi_dev, L, high_bits, n_total, n_zz, pair_bits, angles, 0, nullptr, nullptr,
                      count, bits, thetas, nullptr);
}

int dq_slice_phase_rx_many_scatter(dq_context* ctx, void* psi_dev, int L, uint64_t high_bits, int n_total, int n_zz,
                                   const int32_t* pair_bits, const double* angles, int count, const int32_t* bits,
                                   const double* thetas, int g, int rank, void* const* peer_recv) {
    DQ_TRY(check_scatter("dq_slice_phase_rx_many_scatter", L, count, g, rank, peer_recv));
    Scatter sc{g, rank, peer_recv};
    return slice_step("dq_slice_phase_rx_many_scatter", ctx, psi_dev, L, high_bits, n_total, n_zz, pair_bits, angles, 0, nullptr,
                      nullptr, count, bits, thetas, &sc);
}

int dq_slice_step(dq_context* ctx, void* psi_dev, int L, uint64_t high_bits, int n_total, int n_zz, const int32_t* pair_bits,
                  const double* angles, int n_pre, const int32_t* pre_bits, const double* pre_thetas, int count,
                  const int32_t* bits, const double* thetas) {
    return slice_step("dq_slice_step", ctx, psi_dev, L, high_bits, n_total, n_zz, pair_bits, angles, n_pre, pre_bits, pre_thetas,
                      count, bits, thetas, nullptr);
}

int dq_slice_step_scatter(dq_context* ctx, void* psi_dev, int L, uint64_t high_bits, int n_total, int n_zz,
                          const int32_t* pair_bits, const double* angles, int n_pre, const int32_t* pre_bits,
                          const double* pre_thetas, int count, const int32_t* bits, const double* thetas, int g, int rank,
                          void* const* peer_recv) {
    DQ_TRY(check_scatter("dq_slice_step_scatter", L, count, g, rank, peer_recv));
    Scatter sc{g, rank, peer_recv};
    return slice_step("dq_slice_step_scatter", ctx, psi_dev, L, high_bits, n_total, n_zz, pair_bits, angles, n_pre, pre_bits,
                      pre_thetas, count, bits, thetas, &sc);
}

int dq_slice_evolve_steps(dq_context* ctx, void* psi_dev, int L, uint64_t high_bits, int n_total, int n_zz,
                          const int32_t* pair_bits, int count, const int32_t* bits, int n_steps, const double* angles,
                          int64_t ld_angles, const double* thetas, int64_t ld_thetas) {
    const char* what = "dq_slice_evolve_steps";
    DQ_REQUIRE((g_rec || (ctx && psi_dev)) && (n_steps == 0 || (angles && (count == 0 || thetas))), "NULL argument");
    DQ_REQUIRE(L >= 1 && L <= 33 && n_total >= L && n_total <= 40 && n_steps >= 0, "%s: L=%d n=%d steps=%d", what, L, n_total, n_steps);
    DQ_REQUIRE((high_bits >> (n_total - L)) == 0, "%s: high_bits do not fit %d global bits", what, n_total - L);
    DQ_REQUIRE(ld_angles >= 1 + n_zz && ld_thetas >= count, "%s: row strides %lld, %lld", what, (long long)ld_angles, (long long)ld_thetas);
    if (n_steps == 0) return DQ_OK;
    if (!g_rec) DQ_TRY(ctx->set_device());
    std::vector<std::vector<Rot>> rot(n_steps);
    for (int k = 0; k < n_steps; ++k) DQ_TRY(sorted_rots(what, L, count, bits, thetas + (size_t)k * ld_thetas, rot[k]));
    std::vector<int> tb(rot[0].size());
    for (size_t i = 0; i < rot[0].size(); ++i) tb[i] = rot[0][i].bit;
    const std::vector<TileSet> sets = plan_sets(L, tb);
    PhaseArgs h;
    DQ_TRY(fill_args(h, n_total, n_zz, pair_bits, angles + 1, angles[0]));
    // Chained form: step p leaves the rotations of ONE tile set (d_p) undone; the first pass of step p + 1 runs on that set and
    // applies [d_p's rotations of step p] [phase of step p + 1] [d_p's rotations of step p + 1] -- a step costs one pass less
    // (two tile sets: ONE pass per step).  Two sets that can carry the phase take turns as d_p.
    // (the two with the fewest targets: the boundary pass rotates its targets twice)
    int b0 = -1, b1 = -1;
    for (size_t k = 0; k < sets.size(); ++k)
        if (can_carry_phase(h, sets[k], L, true, false)) {
            if (b0 < 0 || sets[k].n_targets < sets[b0].n_targets) { b1 = b0; b0 = (int)k; }
            else if (b1 < 0 || sets[k].n_targets < sets[b1].n_targets) b1 = (int)k;
        }
    const bool chain = b1 >= 0 && n_steps >= 2 && !getenv("DQ_SLICE_NO_CHAIN");
    const std::vector<Rot> none;
    int owed = -1;                                                    // tile set whose rotations of the previous step are still owed
    for (int k = 0; k < n_steps; ++k) {
        const double* row = angles + (size_t)k * ld_angles;
        DQ_TRY(fill_args(h, n_total, n_zz, pair_bits, row + 1, row[0]));
        int first = -1, skip = -1;
        if (g_rec) g_rec->step = k;
        if (chain) {
            first = owed >= 0 ? owed : b0;
            skip = k + 1 < n_steps ? (first == b0 ? b1 : b0) : -1;
        }
        DQ_TRY(step_impl(ctx, psi_dev, L, high_bits, &h, owed >= 0 ? rots_of(sets[owed], rot[k - 1]) : none, rot[k], sets, first, skip,
                         nullptr, n_total, n_zz, pair_bits, row));
        owed = skip;
    }
    return DQ_OK;
}

int dq_slice_plan(int L, int n_total, int n_zz, const int32_t* pair_bits, int count, const int32_t* bits, int n_steps,
                  int assume_tma, int32_t* rows_out, int64_t rows_cap, int64_t* n_rows_out) {
    DQ_REQUIRE(n_rows_out && (rows_cap == 0 || rows_out), "NULL argument");
    DQ_REQUIRE(n_steps >= 0 && n_steps <= 4096 && count >= 0 && count <= 64 && n_zz >= 0 && n_zz <= kMaxPairs, "dq_slice_plan: steps=%d count=%d n_zz=%d", n_steps, count, n_zz);
    std::vector<double> angles((size_t)std::max(1, n_steps) * (1 + n_zz), 0.01), thetas((size_t)std::max(1, n_steps) * std::max(1, count), 0.1);
    PlanRec rec;
    rec.assume_tma = assume_tma != 0;
    rec.step = 0;
    g_rec = &rec;
    const int st = dq_slice_evolve_steps(nullptr, nullptr, L, 0, n_total, n_zz, pair_bits, count, bits, n_steps, angles.data(), 1 + n_zz,
                                         thetas.data(), std::max(1, count));
    g_rec = nullptr;
    if (st != DQ_OK) return st;
    const int64_t n_rows = (int64_t)rec.rows.size() / 9;
    *n_rows_out = n_rows;
    for (int64_t i = 0; i < std::min(n_rows, rows_cap) * 9; ++i) rows_out[i] = rec.rows[(size_t)i];
    return DQ_OK;
}

int dq_slice_plan_step(int L, int n_total, int n_zz, const int32_t* pair_bits, int n_pre, const int32_t* pre_bits, int count,
                       const int32_t* bits, int scatter_g, int assume_tma, int32_t* rows_out, int64_t rows_cap, int64_t* n_rows_out) {
    DQ_REQUIRE(n_rows_out && (rows_cap == 0 || rows_out), "NULL argument");
    DQ_REQUIRE(n_pre >= 0 && n_pre <= 64 && count >= 0 && count <= 64 && n_zz >= 0 && n_zz <= kMaxPairs && scatter_g >= 0 &&
               (1 << scatter_g) <= kMaxPeers, "dq_slice_plan_step: n_pre=%d count=%d n_zz=%d g=%d", n_pre, count, n_zz, scatter_g);
    std::vector<double> angles((size_t)1 + n_zz, 0.01), thetas((size_t)std::max(1, count), 0.1), pre_thetas((size_t)std::max(1, n_pre), 0.2);
    void* peers[kMaxPeers];
    for (int j = 0; j < kMaxPeers; ++j) peers[j] = &peers[0];          // never dereferenced
    Scatter sc{scatter_g, 0, peers};
    PlanRec rec;
    rec.assume_tma = assume_tma != 0;
    rec.step = 0;
    g_rec = &rec;
    const int st = slice_step("dq_slice_plan_step", nullptr, nullptr, L, 0, n_total, n_zz, pair_bits, angles.data(), n_pre, pre_bits,
                              pre_thetas.data(), count, bits, thetas.data(), scatter_g > 0 ? &sc : nullptr);
    g_rec = nullptr;
    if (st != DQ_OK) return st;
    const int64_t n_rows = (int64_t)rec.rows.size() / 9;
    *n_rows_out = n_rows;
    for (int64_t i = 0; i < std::min(n_rows, rows_cap) * 9; ++i) rows_out[i] = rec.rows[(size_t)i];
    return DQ_OK;
}

int dq_slice_energy(dq_context* ctx, const void* psi_dev, int L, uint64_t high_bits, int n_total, int n_zz,
                    const int32_t* pair_bits, const double* m_zz, double m_const, double* partial_out) {
    DQ_REQUIRE(ctx && psi_dev && partial_out, "NULL argument");
    DQ_REQUIRE(L >= 1 && L <= 33 && n_total >= L && n_total <= 40, "dq_slice_energy: L=%d n=%d", L, n_total);
    DQ_TRY(ctx->set_device());
    PhaseArgs h;
    DQ_TRY(fill_args(h, n_total, n_zz, pair_bits, m_zz, m_const));
    PhaseArgs* d;
    DQ_TRY(upload_args(ctx, h, &d));
    const int blocks = std::min(grid_for(ctx, (size_t)1 << std::max(0, L - kGrayBits)), 1024);
    if (!ctx->slice_partials) DQ_CUDA(cudaMalloc(&ctx->slice_partials, 1024 * sizeof(double)));   // kept: no malloc / free (device-wide syncs) per energy
    double* d_part = (double*)ctx->slice_partials;
    GrayArgs gh;
    if (build_gray(h, L, true, gh)) {
        GrayArgs* dg = nullptr;
        const int st = upload_gray(ctx, gh, &dg);
        if (st != DQ_OK) return st;
        k_slice_energy_gray<<<blocks, kThreads, 0, ctx->stream>>>((const double2*)psi_dev, L, high_bits, d, dg, d_part);
    } else {
        k_slice_energy<<<blocks, kThreads, 0, ctx->stream>>>((const double2*)psi_dev, L, high_bits, d, d_part);
    }
    ctx->launches++;
    std::vector<double> part(blocks);
    cudaError_t e = cudaMemcpyAsync(part.data(), d_part, blocks * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    DQ_CUDA(e);
    double acc = 0.0;
    for (double v : part) acc += v;            // fixed order: deterministic
    *partial_out = acc;
    return DQ_OK;
}

}  // extern "C"
