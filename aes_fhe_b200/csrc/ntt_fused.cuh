// ntt_fused.cuh -- single-launch negacyclic NTT / iNTT: every limb is read from HBM once and
// written once (the algorithmic 2*N*8 bytes), instead of once per pass.
//
// Why: the two-pass transform of ntt.cuh moves 4*N*8 bytes per limb through HBM (the lazy
// intermediate goes out and comes back), which caps it at 50 % of the roofline before any
// arithmetic is counted.  Here the two passes run inside ONE persistent cooperative kernel:
//
//   * CTAs have TPB = 128 threads and own TPB*16 = 2048 elements of a limb per phase; the grid
//     is (resident CTAs per SM, 6) x 148 SMs, cut into groups of GS = N/2048 CTAs (32 for
//     N = 2^16).  A group owns one limb at a time; CTA `rank` does column tile `rank` of the
//     first pass and row tile `rank` of the second.  The hand-over between the passes is a
//     release/acquire counter per group in global memory, software-pipelined by one phase
//     (see the schedule below) so that it is almost never waited on.
//   * the lazy intermediate lives in a per-group, double-buffered scratch limb (2 x 512 KiB x
//     27 groups = 27 MiB) that is rewritten for every limb and therefore never leaves the
//     126 MB L2 (st.global.cg / ld.global.cg): measured DRAM traffic is 1.18 MB per limb.
//   * a group walks its rows limb-major, so consecutive rows share the modulus: the twiddles a
//     row tile needs for its eight stages (and the <= 256 of the column pass) are staged in
//     shared memory once per modulus change and read with LDS, not re-fetched from L2 for
//     every row (the two-pass kernel reads 512 KiB of twiddles per limb from L2).
//   * the row phase is warp-local (a warp owns two 256-element rows: __syncwarp only), the
//     input tile of the next row is prefetched into L2 one phase ahead.
//
// Arithmetic, butterfly order and the load/store functors are those of ntt.cuh, so the
// canonical output is bit-identical to the two-pass kernels and to the oracle.
#pragma once
#include "ntt.cuh"
#include "ntt_chained.cuh"

#ifndef FHE_FUSED_TPB
#define FHE_FUSED_TPB 128
#endif
#ifndef FHE_FUSED_MINB
#define FHE_FUSED_MINB (768 / FHE_FUSED_TPB)
#endif
#define FHE_FUSED_MAX_OCC (1024 / FHE_FUSED_TPB)
#define FHE_FUSED_ROWS (FHE_FUSED_TPB / 16)                       /* 256-element rows per row tile */
#define FHE_FUSED_TILE (FHE_FUSED_TPB * 16)                       /* elements per CTA and phase */
#define FHE_FUSED_XCHG (FHE_FUSED_TILE + FHE_FUSED_TILE / 8)          /* exchange buffer, padded */
#define FHE_FUSED_SMEM_DOUBLES (FHE_FUSED_XCHG + FHE_FUSED_TILE + 256)
#define FHE_FUSED_SMEM_BYTES (FHE_FUSED_SMEM_DOUBLES * 8)
#define FHE_FUSED_SPIN_LIMIT (1u << 24)
#define FHE_FUSED_MAX_WGROUPS 64

struct FusedCtl {
    u64* scratch;          // [groups][2][N] lazy doubles
    unsigned* ctr;         // [groups][4] phase counters, zero at launch
    unsigned* err;
    int groups;
    int flags;                  // 1 = spread a group's CTAs over the age slots of many SMs (see FusedHost::weight)
    int weighted;               // start[] holds the row range of every group (else equal chunks)
    int start[FHE_FUSED_MAX_WGROUPS + 1];
    unsigned long long* gtime;  // [groups] wall time of each group's rank-0 CTA in ns (calibration), or null
    unsigned long long* prof;   // FHE_FUSED_PROFILE builds only
};

#ifndef FHE_EMU
FHE_D u64 ld_cg(const u64* p) { return __ldcg(p); }
FHE_D void st_cg(u64* p, u64 v) { __stcg(p, v); }
FHE_D unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
FHE_D void red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
FHE_D void red_relaxed_add(unsigned* p, unsigned v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
FHE_D void spin_pause() { __nanosleep(32); }
FHE_D void sync_warp() { __syncwarp(); }
#else
#include <thread>
inline u64 ld_cg(const u64* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
inline void st_cg(u64* p, u64 v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
inline unsigned ld_acquire_u32(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void red_release_add(unsigned* p, unsigned v) { __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
inline void red_relaxed_add(unsigned* p, unsigned v) { __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
inline void spin_pause() { std::this_thread::yield(); }
inline void sync_warp() { __syncthreads(); }       // control flow is CTA-uniform, so this is equivalent
#endif

#ifdef FHE_FUSED_PROFILE
#define FHE_PROF_T0 const long long prof_t0 = clock64();
#define FHE_PROF_ADD(var) var += clock64() - prof_t0;
#else
#define FHE_PROF_T0
#define FHE_PROF_ADD(var)
#endif

// Group-scope signalling through global counters.  A warp signals after a warp-level barrier
// (its lanes' global writes / consumed reads are ordered before lane 0's release); a CTA waits
// with thread 0 acquiring, then a CTA barrier.  Counters count WARPS.
FHE_D void warp_signal(unsigned* ctr) {
    sync_warp();
    if ((threadIdx.x & 31) == 0) red_release_add(ctr, 1u);
}
// consumed-reads signal: the loads have returned (their values were used), nothing to publish
FHE_D void warp_signal_relaxed(unsigned* ctr) {
    sync_warp();
    if ((threadIdx.x & 31) == 0) red_relaxed_add(ctr, 1u);
}
FHE_D void group_wait(const unsigned* ctr, unsigned target, unsigned* err) {
    if (threadIdx.x == 0) {
        unsigned spins = 0;
        while (ld_acquire_u32(ctr) < target) {
            spin_pause();
            if (++spins > FHE_FUSED_SPIN_LIMIT) { *err = 1u; break; }     // never hang the GPU
        }
    }
    __syncthreads();
}

// Twiddles of one modulus for row tile `rank` (ROWS rows of 256 elements): stage LOG_R + s
// (s = 1..8) has ROWS * 2^(s-1) distinct twiddles over the tile, stored at offset
// ROWS * (2^(s-1) - 1).
template <int LOG_R>
FHE_D void stage_twiddles(double* s_tw, const double* tw, int rank) {
    constexpr int ROWS = FHE_FUSED_ROWS;
#pragma unroll 1
    for (int s = 1; s <= 8; ++s) {
        const int len = ROWS << (s - 1);
        const double* src = tw + ((size_t)1 << (LOG_R + s - 1)) + ((u32)(rank * FHE_FUSED_TILE) >> (9 - s));
        double* dst = s_tw + (len - ROWS);
        for (int i = threadIdx.x; i < len; i += FHE_FUSED_TPB) dst[i] = ld_d(src + i);
    }
}

// the 15 twiddles of stages LOG_R+5 .. LOG_R+8 for thread tid's 16 contiguous elements
FHE_D Tw15 lds_tw15(const double* s_tw, int tid) {
    constexpr int ROWS = FHE_FUSED_ROWS;
    Tw15 t;
    t.w[0] = s_tw[ROWS * 15 + tid];
#ifndef FHE_EMU
    const double2 a = *reinterpret_cast<const double2*>(s_tw + ROWS * 31 + 2 * tid);
    t.w[1] = a.x; t.w[2] = a.y;
    const double2* p4 = reinterpret_cast<const double2*>(s_tw + ROWS * 63 + 4 * tid);
    const double2 b0 = p4[0], b1 = p4[1];
    t.w[3] = b0.x; t.w[4] = b0.y; t.w[5] = b1.x; t.w[6] = b1.y;
    const double2* p8 = reinterpret_cast<const double2*>(s_tw + ROWS * 127 + 8 * tid);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const double2 c = p8[j]; t.w[7 + 2 * j] = c.x; t.w[8 + 2 * j] = c.y; }
#else
    for (int j = 0; j < 2; ++j) t.w[1 + j] = s_tw[ROWS * 31 + 2 * tid + j];
    for (int j = 0; j < 4; ++j) t.w[3 + j] = s_tw[ROWS * 63 + 4 * tid + j];
    for (int j = 0; j < 8; ++j) t.w[7 + j] = s_tw[ROWS * 127 + 8 * tid + j];
#endif
    return t;
}
// twiddle of stage LOG_R + s (s = 1..4) for local element e of the row tile
FHE_D double lds_tw_lo(const double* s_tw, int s, int e) {
    return s_tw[FHE_FUSED_ROWS * ((1 << (s - 1)) - 1) + (e >> (9 - s))];
}

// Schedule of one CTA over the rows r_0, r_1, ... of its group (software-pipelined by one
// phase so that a hand-over is almost never waited on):
//      A(r_0);  for k = 0, 1, ...:  A(r_{k+1});  wait "all A(r_k) stored";  B(r_k)
// Row r_j uses scratch buffer j % 2 and the counters of parity j % 2 (a CTA may be one row
// ahead of another, never two, so arrivals for r_j and r_{j+2} cannot mix; r_j and r_{j+1} use
// different counters).  Every warp signals ctrA[j%2] after its A(r_j) stores; B(r_j) waits for
// GS * WARPS * (j/2 + 1) arrivals.  Before A(r_j) stores into the buffer, every warp of the
// group must have loaded its B(r_{j-2}) inputs from it: a warp signals ctrB[j%2] once its loads
// are consumed, A waits for GS * WARPS * (j/2) just before storing.
// Walks the launch order y = j * n_blocks + blk (limb-major) incrementally: no divisions per row.
struct FusedCursor {
    int y, y1;
    RowRef r;          // poly block and limb slot of row y
    int dig;           // blk % digits (ModUp layout), else 0
    FHE_D void init(const RowMap& map, int y0, int y_end) {
        y = y0; y1 = y_end;
        r.j = y0 / map.n_blocks; r.blk = y0 - r.j * map.n_blocks;
        dig = map.skip_alpha > 0 ? r.blk % map.digits : 0;
        seek(map);
    }
    FHE_D int mid(const RowMap& map) const { return map.mod_id_of(r.j, dig); }
    FHE_D void step(const RowMap& map) {
        ++y;
        if (++r.blk == map.n_blocks) { r.blk = 0; ++r.j; dig = 0; }
        else if (map.skip_alpha > 0 && ++dig == map.digits) dig = 0;
    }
    FHE_D void seek(const RowMap& map) { while (y < y1 && mid(map) < 0) step(map); }
    FHE_D void next(const RowMap& map) { step(map); seek(map); }
};

// Index arithmetic below is written so that, after unrolling, every shared / global access
// is `base register + compile-time offset` (N, the tile shape and the stage are template
// constants; indices are signed ints so that `base + 16 * i` folds into the address
// immediate).  The non-FP64 instruction count of the transform is what decides whether the
// FP64 pipe or the issue slots saturate first.
//
// Twiddle index algebra (element e of the limb, stage S, LOG_N = LOG_R + 8):
//   tw_index(S, e) = 2^(S-1) + (e >> (LOG_N - S + 1)).
//   column phase, first round  (S = 1..LEV1, registers hold rows g + G i):       2^(S-1) + (i >> (5-S))
//   column phase, second round (S = LEV1+t, t = 1..4, rows 16 g + i):            2^(S-1) + (g << (t-1)) + (i >> (5-t))
//   row phase: stage LOG_R + s uses the staged table, offset ROWS (2^(s-1) - 1) + (local element >> (9-s)).
template <int LOG_R>
struct FusedGeom {
    static constexpr int TPB = FHE_FUSED_TPB, WARPS = TPB / 32, TILE = FHE_FUSED_TILE, ROWS = FHE_FUSED_ROWS;
    static constexpr int R = 1 << LOG_R, G = R / 16, COLS = TPB / G, LEV1 = LOG_R - 4, GS = (R * 256) / TILE;
    static constexpr int LOG_N = LOG_R + 8;
    static constexpr int RSTRIDE = 16 * 18;        // row phase: a 256-element row as 16 blocks of 16 (+2 pad: LDS.128, no conflicts)
    static constexpr int PADA = COLS < 16 ? COLS : 0;   // column phase: pad per 16 rows so a half-warp's rows hit distinct banks
};

// column-phase twiddle, first round: thread-independent
template <int LOG_R>
FHE_D double twa_r1(const double* s_twa, int S, int i) { return s_twa[(1 << (S - 1)) + (i >> (5 - S))]; }
// column-phase twiddle, second round: s_twa_g = s_twa + (g << (t-1)) is formed by the caller per level
template <int LOG_R>
FHE_D double twa_r2(const double* s_twa, int g, int S, int i) {
    constexpr int LEV1 = LOG_R - 4;
    const int t = S - LEV1;
    return s_twa[(1 << (S - 1)) + (g << (t - 1)) + (i >> (5 - t))];
}

// ------------------------------------------------------------------ forward
template <int LOG_R, class LoadOp, class StoreOp>
__global__ void __launch_bounds__(FHE_FUSED_TPB, FHE_FUSED_MINB) ntt_fwd_fused(DevTables T, RowMap map, int rows, LoadOp ld,
                                                                     StoreOp st, FusedCtl fc) {
    typedef FusedGeom<LOG_R> GM;
    constexpr int TPB = GM::TPB, WARPS = GM::WARPS, TILE = GM::TILE, ROWS = GM::ROWS, R = GM::R, G = GM::G,
                  COLS = GM::COLS, LEV1 = GM::LEV1, GS = GM::GS, LOG_N = GM::LOG_N, RS = GM::RSTRIDE, PADA = GM::PADA;
    FHE_DYN_SHARED(double, smem);
    double* sm = smem;                          // [FHE_FUSED_XCHG] exchange
    double* s_tw = smem + FHE_FUSED_XCHG;        // [TILE - ROWS] row-tile twiddles
    double* s_twa = s_tw + TILE;                // [R <= 256]   column-pass twiddles
    // Group = GS consecutive CTAs.  The hardware fills the SMs round-robin in blockIdx order, so a
    // group's CTAs sit on GS different SMs in the SAME residency slot, and the CTAs sharing an SM
    // belong to different groups (one group's hand-over wait is another's FP64 time).  The warp
    // scheduler favours older CTAs, so groups in later slots run slower: the host gives every group
    // a row range proportional to its calibrated speed (fc.start).
    const int grp = (fc.flags & 1) ? blockIdx.x % fc.groups : blockIdx.x / GS;
    const int rank = (fc.flags & 1) ? blockIdx.x / fc.groups : blockIdx.x % GS;
    const int chunk = (rows + fc.groups - 1) / fc.groups;
    const int y_lo = fc.weighted ? fc.start[grp] : grp * chunk;
    const int y_hi = fc.weighted ? fc.start[grp + 1] : min(rows, grp * chunk + chunk);
#ifndef FHE_EMU
    unsigned long long gt0 = 0;
    if (fc.gtime) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
#endif
    const int tid = threadIdx.x;
    unsigned* ctrA = fc.ctr + 32 * grp;         // [2] by row parity: "first phase of row j stored"
    unsigned* ctrB = ctrA + 2;                  // [2] by row parity: "second phase of row j has consumed its inputs"
    u64* scratch = fc.scratch + ((size_t)(2 * grp) << LOG_N);
    FusedCursor ca, cb;
    ca.init(map, y_lo, y_hi);
    cb = ca;
    unsigned na = 0, nb = 0;
    int mid_a = -1, mid_b = -1;
    // thread geometry of the two phases
    const int cc = tid % COLS, g = tid / COLS;          // column phase: column cc of the tile, row residue g
    const int c = rank * COLS + cc;
    const int l16 = tid & 15, rr = tid >> 4;            // row phase: 16 threads per 256-element row
    const int rbase = ((rank * ROWS + rr) << 8) + l16;  // limb index of x[0] in the strided view
    double* smS = sm + rr * RS + l16;                   // strided view    [18 * i]
    double* smC = sm + rr * RS + 18 * l16;              // contiguous view [i], 16-byte aligned
#ifdef FHE_FUSED_PROFILE
    long long prof_wa = 0, prof_wb = 0, prof_pa = 0, prof_pb = 0;
    const long long prof_start = clock64();
    unsigned long long prof_gt0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prof_gt0));
#endif

    auto phase_a = [&]() {   // column tile `rank` of row ca: stages 1 .. LOG_R, lazy result to scratch
        const RowRef row = ca.r;
        const int mid = ca.mid(map);
        __syncthreads();
        if (mid != mid_a) {
            const double* tw = T.tw_fwd + ((size_t)mid << LOG_N);
            for (int i = tid; i < R; i += TPB) s_twa[i] = ld_d(tw + i);
            mid_a = mid;
            __syncthreads();
        }
        const Modulus M = T.mod[mid];
        const double q = M.qd, qinv = M.qinv;
        u64* scr = scratch + ((size_t)(na & 1) << LOG_N);
        double x[16];
        if (LEV1 > 0) {
            const int e0 = (g << 8) + c;
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = ld(map, row, e0 + i * (G << 8), mid, M);
            ct_radix16<(16 >> (LEV1 > 0 ? LEV1 : 1))>(x, 1, [&](int S, int, int i) {
                return mk_tw(twa_r1<LOG_R>(s_twa, S, i), qinv); }, q);
            // row r of the tile lives at r * COLS + (r >> 4) * PADA
            double* smw = sm + g * COLS + cc;
#pragma unroll
            for (int i = 0; i < 16; ++i) smw[i * (G * COLS) + (i >> (4 - LEV1)) * PADA] = x[i];
            __syncthreads();
            const double* smr = sm + g * (16 * COLS + PADA) + cc;
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = smr[i * COLS];
        } else {
            const int e0 = (g << 12) + c;
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = ld(map, row, e0 + (i << 8), mid, M);
        }
        ct_radix16<1>(x, LEV1 + 1, [&](int S, int, int i) { return mk_tw(twa_r2<LOG_R>(s_twa, g, S, i), qinv); }, q);
        { FHE_PROF_T0 if (na >= 2) group_wait(ctrB + (na & 1), GS * WARPS * (na >> 1), fc.err); FHE_PROF_ADD(prof_wb) }
        u64* so = scr + (g << 12) + c;
#pragma unroll
        for (int i = 0; i < 16; ++i) st_cg(so + (i << 8), f_to_bits(x[i]));
        warp_signal(ctrA + (na & 1));
        ++na; ca.next(map);
        if (ca.y < ca.y1)      // input tile of the next column phase -> L2, one phase ahead
            for (int r = tid; r < R; r += TPB) ld.prefetch(map, ca.r, (r << 8) + rank * COLS, COLS);
    };

    auto phase_b = [&]() {   // row tile `rank` of row cb: stages LOG_R+1 .. LOG_N on ROWS contiguous rows of 256
        const RowRef row = cb.r;
        const int mid = cb.mid(map);
        __syncthreads();
        if (mid != mid_b) {
            stage_twiddles<LOG_R>(s_tw, T.tw_fwd + ((size_t)mid << LOG_N), rank);
            mid_b = mid;
        }
        { FHE_PROF_T0 group_wait(ctrA + (nb & 1), GS * WARPS * ((nb >> 1) + 1), fc.err); FHE_PROF_ADD(prof_wa) }
        const Modulus M = T.mod[mid];
        const double q = M.qd, qinv = M.qinv;
        const u64* scr = scratch + ((size_t)(nb & 1) << LOG_N) + rbase;
        double x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = bits_to_f(ld_cg(scr + 16 * i));
        {   // stages LOG_R+1 .. LOG_R+4: twiddle of level s for register group starting at i
            const double* t1 = s_tw + rr, *t2 = s_tw + ROWS + 2 * rr, *t3 = s_tw + 3 * ROWS + 4 * rr, *t4 = s_tw + 7 * ROWS + 8 * rr;
            ct_radix16<1>(x, 1, [&](int s, int, int i) {
                return mk_tw(s == 1 ? t1[0] : s == 2 ? t2[i >> 3] : s == 3 ? t3[i >> 2] : t4[i >> 1], qinv); }, q);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) smS[18 * i] = x[i];
        warp_signal_relaxed(ctrB + (nb & 1));                      // this warp's scratch reads are consumed
        lds16(x, smC);
        const Tw15 t2 = lds_tw15(s_tw, tid);
        ct_radix16<1>(x, 5, [&](int, int half, int i) { return mk_tw(t2.w[8 / half - 1 + i / (2 * half)], qinv); }, q);
        sync_warp();
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = reduce_canon(x[i], q, qinv);
        sts16(smC, x);
        sync_warp();
#pragma unroll
        for (int i = 0; i < 16; ++i) st(map, row, rbase + 16 * i, smS[18 * i], mid, M);
        ++nb; cb.next(map);
    };

    int pending = 0;
    while (ca.y < ca.y1 || pending > 0) {
        if (ca.y < ca.y1) { FHE_PROF_T0 phase_a(); ++pending; FHE_PROF_ADD(prof_pa) }
        if (pending == 2 || !(ca.y < ca.y1)) { FHE_PROF_T0 phase_b(); --pending; FHE_PROF_ADD(prof_pb) }
    }
#ifndef FHE_EMU
    if (fc.gtime && rank == 0 && tid == 0) {
        unsigned long long gt1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
        fc.gtime[grp] = gt1 - gt0;
    }
#endif
#ifdef FHE_FUSED_PROFILE
    if (tid == 0 && fc.prof) {
        atomicAdd(fc.prof + 0, (unsigned long long)(clock64() - prof_start));
        atomicAdd(fc.prof + 1, (unsigned long long)prof_wa);
        atomicAdd(fc.prof + 2, (unsigned long long)prof_wb);
        atomicAdd(fc.prof + 3, (unsigned long long)prof_pa);
        atomicAdd(fc.prof + 4, (unsigned long long)prof_pb);
        atomicAdd(fc.prof + 5, (unsigned long long)nb);
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        unsigned long long* rec = fc.prof + 8 + 4 * (size_t)blockIdx.x;
        rec[0] = smid; rec[1] = prof_gt0; rec[2] = t1; rec[3] = (unsigned long long)(clock64() - prof_start);
    }
#endif
}

// ------------------------------------------------------------------ inverse
template <int LOG_R, class LoadOp, class StoreOp>
__global__ void __launch_bounds__(FHE_FUSED_TPB, FHE_FUSED_MINB) ntt_inv_fused(DevTables T, RowMap map, int rows, LoadOp ld,
                                                                     StoreOp st, const ConstF* scale, FusedCtl fc) {
    typedef FusedGeom<LOG_R> GM;
    constexpr int TPB = GM::TPB, WARPS = GM::WARPS, TILE = GM::TILE, ROWS = GM::ROWS, R = GM::R, G = GM::G,
                  COLS = GM::COLS, LEV1 = GM::LEV1, GS = GM::GS, LOG_N = GM::LOG_N, RS = GM::RSTRIDE, PADA = GM::PADA;
    FHE_DYN_SHARED(double, smem);
    double* sm = smem;
    double* s_tw = smem + FHE_FUSED_XCHG;
    double* s_twa = s_tw + TILE;
    // Group = GS consecutive CTAs.  The hardware fills the SMs round-robin in blockIdx order, so a
    // group's CTAs sit on GS different SMs in the SAME residency slot, and the CTAs sharing an SM
    // belong to different groups (one group's hand-over wait is another's FP64 time).  The warp
    // scheduler favours older CTAs, so groups in later slots run slower: the host gives every group
    // a row range proportional to its calibrated speed (fc.start).
    const int grp = (fc.flags & 1) ? blockIdx.x % fc.groups : blockIdx.x / GS;
    const int rank = (fc.flags & 1) ? blockIdx.x / fc.groups : blockIdx.x % GS;
    const int chunk = (rows + fc.groups - 1) / fc.groups;
    const int y_lo = fc.weighted ? fc.start[grp] : grp * chunk;
    const int y_hi = fc.weighted ? fc.start[grp + 1] : min(rows, grp * chunk + chunk);
#ifndef FHE_EMU
    unsigned long long gt0 = 0;
    if (fc.gtime) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
#endif
    const int tid = threadIdx.x;
    unsigned* ctrA = fc.ctr + 32 * grp;
    unsigned* ctrB = ctrA + 2;
    u64* scratch = fc.scratch + ((size_t)(2 * grp) << LOG_N);
    FusedCursor ca, cb;
    ca.init(map, y_lo, y_hi);
    cb = ca;
    unsigned na = 0, nb = 0;
    int mid_a = -1, mid_b = -1;
    ConstF dummy; dummy.w = 0; dummy.wq = 0;
    const int cc = tid % COLS, g = tid / COLS;
    const int c = rank * COLS + cc;
    const int l16 = tid & 15, rr = tid >> 4;
    const int rbase = ((rank * ROWS + rr) << 8) + l16;
    double* smS = sm + rr * RS + l16;
    double* smC = sm + rr * RS + 18 * l16;

    auto phase_a = [&]() {   // row tile `rank` of row ca: stages LOG_N .. LOG_R+1 (warp-local)
        const RowRef row = ca.r;
        const int mid = ca.mid(map);
        __syncthreads();
        if (mid != mid_a) {
            stage_twiddles<LOG_R>(s_tw, T.tw_inv + ((size_t)mid << LOG_N), rank);
            mid_a = mid;
            __syncthreads();
        }
        const Modulus M = T.mod[mid];
        const double q = M.qd, qinv = M.qinv;
        u64* scr = scratch + ((size_t)(na & 1) << LOG_N) + rbase;
        double x[16];
        const Tw15 t1 = lds_tw15(s_tw, tid);
#pragma unroll
        for (int i = 0; i < 16; ++i) smS[18 * i] = ld(map, row, rbase + 16 * i, mid, M);
        sync_warp();
        lds16(x, smC);
        gs_radix16<1, false>(x, 8, [&](int, int half, int i) { return mk_tw(t1.w[8 / half - 1 + i / (2 * half)], qinv); },
                             q, dummy, dummy);
        sync_warp();
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = reduce_sym(x[i], q, qinv);
        sts16(smC, x);
        sync_warp();
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = smS[18 * i];
        {
            const double* t1p = s_tw + rr, *t2p = s_tw + ROWS + 2 * rr, *t3p = s_tw + 3 * ROWS + 4 * rr, *t4p = s_tw + 7 * ROWS + 8 * rr;
            gs_radix16<1, false>(x, 4, [&](int s, int, int i) {
                return mk_tw(s == 1 ? t1p[0] : s == 2 ? t2p[i >> 3] : s == 3 ? t3p[i >> 2] : t4p[i >> 1], qinv); }, q, dummy, dummy);
        }
        if (na >= 2) group_wait(ctrB + (na & 1), GS * WARPS * (na >> 1), fc.err);
#pragma unroll
        for (int i = 0; i < 16; ++i) st_cg(scr + 16 * i, f_to_bits(reduce_sym(x[i], q, qinv)));
        warp_signal(ctrA + (na & 1));
        ++na; ca.next(map);
        if (ca.y < ca.y1)      // input rows of the next row phase -> L2, one phase ahead
            ld.prefetch(map, ca.r, ((rank * ROWS) << 8) + 16 * tid, 16);
    };

    auto phase_b = [&]() {   // column tile `rank` of row cb: stages LOG_R .. 1, the last carries the scaling
        const RowRef row = cb.r;
        const int mid = cb.mid(map);
        __syncthreads();
        if (mid != mid_b) {
            const double* tw = T.tw_inv + ((size_t)mid << LOG_N);
            for (int i = tid; i < R; i += TPB) s_twa[i] = ld_d(tw + i);
            mid_b = mid;
        }
        group_wait(ctrA + (nb & 1), GS * WARPS * ((nb >> 1) + 1), fc.err);
        const Modulus M = T.mod[mid];
        const double q = M.qd, qinv = M.qinv;
        const u64* scr = scratch + ((size_t)(nb & 1) << LOG_N) + (g << 12) + c;
        const ConstF* fin = scale ? scale + 2 * (size_t)row.j : T.inv_final + 2 * (size_t)mid;
        const ConstF fin0 = fin[0], fin1 = fin[1];
        double x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = bits_to_f(ld_cg(scr + (i << 8)));
        if (LEV1 > 0) {
            gs_radix16<1, false>(x, LOG_R, [&](int S, int, int i) { return mk_tw(twa_r2<LOG_R>(s_twa, g, S, i), qinv); }, q, fin0, fin1);
            double* smw = sm + g * (16 * COLS + PADA) + cc;
#pragma unroll
            for (int i = 0; i < 16; ++i) smw[i * COLS] = reduce_sym(x[i], q, qinv);
            warp_signal_relaxed(ctrB + (nb & 1));
            __syncthreads();
            const double* smr = sm + g * COLS + cc;
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = smr[i * (G * COLS) + (i >> (4 - LEV1)) * PADA];
            gs_radix16<(16 >> (LEV1 > 0 ? LEV1 : 1)), true>(x, LEV1, [&](int S, int, int i) {
                return mk_tw(twa_r1<LOG_R>(s_twa, S, i), qinv); }, q, fin0, fin1);
            const int e0 = (g << 8) + c;
#pragma unroll
            for (int i = 0; i < 16; ++i) st(map, row, e0 + i * (G << 8), x[i], mid, M);
        } else {
            gs_radix16<1, true>(x, LOG_R, [&](int S, int, int i) { return mk_tw(twa_r2<LOG_R>(s_twa, g, S, i), qinv); }, q, fin0, fin1);
            warp_signal_relaxed(ctrB + (nb & 1));
            const int e0 = (g << 12) + c;
#pragma unroll
            for (int i = 0; i < 16; ++i) st(map, row, e0 + (i << 8), x[i], mid, M);
        }
        ++nb; cb.next(map);
    };

    int pending = 0;
    while (ca.y < ca.y1 || pending > 0) {
        if (ca.y < ca.y1) { phase_a(); ++pending; }
        if (pending == 2 || !(ca.y < ca.y1)) { phase_b(); --pending; }
    }
#ifndef FHE_EMU
    if (fc.gtime && rank == 0 && tid == 0) {
        unsigned long long gt1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
        fc.gtime[grp] = gt1 - gt0;
    }
#endif
}

// ------------------------------------------------------------------ host side
struct FusedHost {
    u64* scratch = nullptr;        // [max_groups][2][N]
    unsigned* ctr = nullptr;       // [4 * max_groups + 1], the last word is the error flag
    unsigned long long* prof = nullptr;
    int max_groups = 0;
    int sm_count = 0;
    int enabled = 0;
    int flags = 0;
    // calibrated relative speed of every group of the full grid (rows per unit time); cal_groups = 0: none
    double weight[FHE_FUSED_MAX_WGROUPS] = {};
    int cal_groups = 0;
    unsigned long long* gtime = nullptr;      // [FHE_FUSED_MAX_WGROUPS] device, filled while calibrating
    int calibrating = 0;
    ChainHost chain;                          // the chained single-launch variant (ntt_chained.cuh)
};

#ifndef FHE_EMU
template <typename... KArgs>
inline int fused_occupancy(void (*k)(KArgs...)) {
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, FHE_FUSED_SMEM_BYTES) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, FHE_FUSED_TPB, FHE_FUSED_SMEM_BYTES) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return occ < FHE_FUSED_MAX_OCC ? occ : FHE_FUSED_MAX_OCC;
}
template <typename... KArgs, typename... Args>
inline bool fused_launch(void (*k)(KArgs...), int ctas, cudaStream_t s, Args... args) {
    void* argv[] = {(void*)&args...};
    return cudaLaunchCooperativeKernel((const void*)k, dim3(ctas), dim3(FHE_FUSED_TPB), argv, FHE_FUSED_SMEM_BYTES, s) == cudaSuccess;
}
#else
void fhe_emu_launch_coresident(dim3 g, dim3 b, const std::function<void()>& body);
template <typename... KArgs>
inline int fused_occupancy(void (*)(KArgs...)) { return 1; }
template <typename... KArgs, typename... Args>
inline bool fused_launch(void (*k)(KArgs...), int ctas, cudaStream_t, Args... args) {
    fhe_emu_set_dyn_smem(FHE_FUSED_SMEM_BYTES);
    fhe_emu_launch_coresident(dim3(ctas), dim3(FHE_FUSED_TPB), [=]() { k(args...); });
    return true;
}
#endif

// how many groups of GS CTAs to launch for `rows` rows; 0 = do not use the fused kernel
inline int fused_groups(const FusedHost& fz, int occ, int gs, int rows) {
    if (!fz.enabled || occ < 1 || rows < 1) return 0;
    int groups = fz.sm_count * occ / gs;
    if (groups < 1) groups = 1;
    if (groups > fz.max_groups) groups = fz.max_groups;
    if (groups > rows) groups = rows;
    if (groups == fz.cal_groups && rows >= 4 * groups) return groups;       // weighted partition of the full grid
    // equalise: the slowest group does ceil(rows / groups) rows; use the fewest groups that achieve it
    const int chunk = (rows + groups - 1) / groups;
    return (rows + chunk - 1) / chunk;
}
// row range of every group: proportional to the calibrated group speeds when the full grid runs
inline void fused_fill_ctl(FusedCtl& fc, const FusedHost& fz, int groups, int rows) {
    fc.scratch = fz.scratch; fc.ctr = fz.ctr; fc.err = fz.ctr + 32 * fz.max_groups; fc.groups = groups;
    fc.prof = fz.prof; fc.flags = fz.flags; fc.weighted = 0; fc.gtime = fz.calibrating ? fz.gtime : nullptr;
    if (groups == fz.cal_groups && groups <= FHE_FUSED_MAX_WGROUPS && rows >= 4 * groups && !(fz.flags & 1)) {
        double tot = 0.0, acc = 0.0;
        for (int g = 0; g < groups; ++g) tot += fz.weight[g];
        fc.start[0] = 0;
        for (int g = 0; g < groups; ++g) {
            acc += fz.weight[g];
            int e = (int)(rows * (acc / tot) + 0.5);
            if (e < fc.start[g]) e = fc.start[g];
            if (e > rows) e = rows;
            fc.start[g + 1] = e;
        }
        fc.start[groups] = rows;
        fc.weighted = 1;
    }
}

template <int LOG_R, class LoadOp, class StoreOp>
inline bool ntt_forward_fused(const DevTables& T, const FusedHost& fz, const RowMap& map, int rows, LoadOp ld,
                              StoreOp st, cudaStream_t s) {
    static const int occ = fused_occupancy(ntt_fwd_fused<LOG_R, LoadOp, StoreOp>);
    constexpr int GS = ((1 << LOG_R) * 256) / FHE_FUSED_TILE;
    const int groups = fused_groups(fz, occ, GS, rows);
    if (groups < 1) return false;
    FusedCtl fc; fused_fill_ctl(fc, fz, groups, rows);
    cudaMemsetAsync(fz.ctr, 0, sizeof(unsigned) * 32 * groups, s);
    return fused_launch(ntt_fwd_fused<LOG_R, LoadOp, StoreOp>, groups * GS, s, T, map, rows, ld, st, fc);
}
template <int LOG_R, class LoadOp, class StoreOp>
inline bool ntt_inverse_fused(const DevTables& T, const FusedHost& fz, const RowMap& map, int rows, LoadOp ld,
                              StoreOp st, const ConstF* scale, cudaStream_t s) {
    static const int occ = fused_occupancy(ntt_inv_fused<LOG_R, LoadOp, StoreOp>);
    constexpr int GS = ((1 << LOG_R) * 256) / FHE_FUSED_TILE;
    const int groups = fused_groups(fz, occ, GS, rows);
    if (groups < 1) return false;
    FusedCtl fc; fused_fill_ctl(fc, fz, groups, rows);
    cudaMemsetAsync(fz.ctr, 0, sizeof(unsigned) * 32 * groups, s);
    return fused_launch(ntt_inv_fused<LOG_R, LoadOp, StoreOp>, groups * GS, s, T, map, rows, ld, st, scale, fc);
}

// Dispatch: the fused single-launch transform when the context enables it, else the two-pass
// kernels of ntt.cuh.  Returns the number of kernels launched.
template <class LoadOp, class StoreOp>
inline int ntt_forward_auto(const DevTables& T, const FusedHost& fz, const RowMap& map_in, int rows, LoadOp ld,
                            u64* work, long long work_stride, StoreOp st, cudaStream_t s) {
    if (fz.chain.enabled && ntt_forward_chained(T, fz.chain, map_in, rows, ld, work, work_stride, st, s)) return 1;
    if (fz.enabled) {
        RowMap map = map_in; map.n_blocks = rows / map.rows_per_poly;
        bool ok = false;
        switch (T.log_n - 8) {
            case 4: ok = ntt_forward_fused<4>(T, fz, map, rows, ld, st, s); break;
            case 5: ok = ntt_forward_fused<5>(T, fz, map, rows, ld, st, s); break;
            case 6: ok = ntt_forward_fused<6>(T, fz, map, rows, ld, st, s); break;
            case 7: ok = ntt_forward_fused<7>(T, fz, map, rows, ld, st, s); break;
            default: ok = ntt_forward_fused<8>(T, fz, map, rows, ld, st, s); break;
        }
        if (ok) return 1;
    }
    ntt_forward(T, map_in, rows, ld, work, work_stride, st, s);
    return 2;
}
template <class LoadOp, class StoreOp>
inline int ntt_inverse_auto(const DevTables& T, const FusedHost& fz, const RowMap& map_in, int rows, LoadOp ld,
                            u64* work, long long work_stride, StoreOp st, const ConstF* scale, cudaStream_t s) {
    if (fz.chain.enabled && ntt_inverse_chained(T, fz.chain, map_in, rows, ld, work, work_stride, st, scale, s)) return 1;
    if (fz.enabled) {
        RowMap map = map_in; map.n_blocks = rows / map.rows_per_poly;
        bool ok = false;
        switch (T.log_n - 8) {
            case 4: ok = ntt_inverse_fused<4>(T, fz, map, rows, ld, st, scale, s); break;
            case 5: ok = ntt_inverse_fused<5>(T, fz, map, rows, ld, st, scale, s); break;
            case 6: ok = ntt_inverse_fused<6>(T, fz, map, rows, ld, st, scale, s); break;
            case 7: ok = ntt_inverse_fused<7>(T, fz, map, rows, ld, st, scale, s); break;
            default: ok = ntt_inverse_fused<8>(T, fz, map, rows, ld, st, scale, s); break;
        }
        if (ok) return 1;
    }
    ntt_inverse(T, map_in, rows, ld, work, work_stride, st, scale, s);
    return 2;
}

// number of groups of the full persistent grid for ring size 2^log_n (plain forward kernel)
inline int fused_full_groups(int log_n, const FusedHost& fz) {
    int occ = 0;
    switch (log_n - 8) {
        case 4: occ = fused_occupancy(ntt_fwd_fused<4, LoadPlain, StorePlain>); break;
        case 5: occ = fused_occupancy(ntt_fwd_fused<5, LoadPlain, StorePlain>); break;
        case 6: occ = fused_occupancy(ntt_fwd_fused<6, LoadPlain, StorePlain>); break;
        case 7: occ = fused_occupancy(ntt_fwd_fused<7, LoadPlain, StorePlain>); break;
        default: occ = fused_occupancy(ntt_fwd_fused<8, LoadPlain, StorePlain>); break;
    }
    const int gs = ((1 << log_n) / FHE_FUSED_TILE) > 0 ? ((1 << log_n) / FHE_FUSED_TILE) : 1;
    int groups = fz.sm_count * occ / gs;
    if (groups < 1) groups = 1;
    if (groups > fz.max_groups) groups = fz.max_groups;
    return groups;
}
