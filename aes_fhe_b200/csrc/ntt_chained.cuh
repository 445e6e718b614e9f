// ntt_chained.cuh -- the two passes of ntt.cuh in ONE launch, ordered so that the lazy
// intermediate is consumed while it is still in the 126 MB L2 (it is written with st.cg, read with
// ld.cg, and overwritten in place by the second pass, so most of it never reaches HBM).
//
// The rows of a launch are cut into chunks of CR limbs (16 MiB of intermediate for CR = 32).
// CTAs take a ticket at start (one atomicAdd) and the ticket order is
//        A(0) A(1) B(0) A(2) B(1) A(3) B(2) ... A(n-1) B(n-2) B(n-1)
// (A(c) = all first-pass tiles of chunk c, B(c) = all second-pass tiles).  A second-pass CTA waits
// until the counter of its chunk says every first-pass CTA of that chunk has stored and released
// its tile; those CTAs hold smaller tickets, i.e. they started earlier, so the wait cannot
// deadlock whatever order the hardware dispatches blocks in, and because a whole segment lies
// between A(c) and B(c) it is almost never a real wait.  Unlike the persistent kernel of
// ntt_fused.cuh the CTAs are short-lived: there is no group to keep in step and no dependence on
// the age priority of the warp scheduler.
#pragma once
#include "ntt.cuh"

#ifndef FHE_CHAIN_MINB
#define FHE_CHAIN_MINB 4
#endif
#define FHE_CHAIN_ROWS 32
#define FHE_CHAIN_MAX_CHUNKS 4096
#define FHE_CHAIN_SPIN_LIMIT (1u << 22)

struct ChainCtl {
    unsigned* ticket;      // [1], zero at launch
    unsigned* done;        // [n_chunks], zero at launch: first-pass CTAs of the chunk that have finished
    unsigned* err;         // set if a wait ever times out (never hangs the GPU)
    int rows, chunk_rows, n_chunks;
};

#ifndef FHE_EMU
FHE_D u64 chain_ld_cg(const u64* p) { return __ldcg(p); }
FHE_D void chain_st_cg(u64* p, u64 v) { __stcg(p, v); }
FHE_D unsigned chain_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
FHE_D void chain_red_release(unsigned* p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(p) : "memory");
}
FHE_D unsigned chain_take_ticket(unsigned* p) { return atomicAdd(p, 1u); }
FHE_D void chain_pause() { __nanosleep(64); }
#else
#include <thread>
inline u64 chain_ld_cg(const u64* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
inline void chain_st_cg(u64* p, u64 v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
inline unsigned chain_ld_acquire(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void chain_red_release(unsigned* p) { __atomic_fetch_add(p, 1u, __ATOMIC_ACQ_REL); }
inline unsigned chain_take_ticket(unsigned* p) { return __atomic_fetch_add(p, 1u, __ATOMIC_ACQ_REL); }
inline void chain_pause() { std::this_thread::yield(); }
#endif

// lazy intermediate through L2 only
struct LoadRawCG {
    const u64* src; long long poly_stride;
    FHE_D double operator()(const RowMap& map, RowRef row, int idx, int, const Modulus&) const {
#ifdef FHE_NTT_DIAG_NOMEM
        return u64_to_f((u64)(idx * 2654435761u + row.j));
#else
        return bits_to_f(chain_ld_cg(src + row_off(map, row, poly_stride) + idx));
#endif
    }
};
struct StoreRawCG {
    u64* dst; long long poly_stride;
    FHE_D void operator()(const RowMap& map, RowRef row, int idx, double v, int, const Modulus&) const {
#ifdef FHE_NTT_DIAG_NOMEM
        if (v == -12345.0) chain_st_cg(dst + row_off(map, row, poly_stride) + idx, f_to_bits(v));
#else
        chain_st_cg(dst + row_off(map, row, poly_stride) + idx, f_to_bits(v));
#endif
    }
};

// ticket -> (second pass?, chunk, position inside the segment)
struct ChainSlot { int second, chunk, pos; };
FHE_D ChainSlot chain_decode(unsigned ticket, int seg, int n) {
    ChainSlot r;
    const int s = (int)(ticket / (unsigned)seg);
    r.pos = (int)(ticket - (unsigned)s * (unsigned)seg);
    if (s == 0) { r.second = 0; r.chunk = 0; }
    else if (s <= 2 * n - 3) {
        if (s & 1) { r.second = 0; r.chunk = (s + 1) >> 1; }
        else { r.second = 1; r.chunk = (s >> 1) - 1; }
    } else { r.second = 1; r.chunk = s - n; }          // s = 2n-2 -> n-2, s = 2n-1 -> n-1
    return r;
}

template <int LOG_R>
FHE_D bool chain_begin(const RowMap& map, const ChainCtl& cc, ChainSlot& sl, RowRef& rref, int& mid, int& tile,
                       unsigned* s_ticket) {
    constexpr int TILES = 1 << (LOG_R - 4);
    if (threadIdx.x == 0) *s_ticket = chain_take_ticket(cc.ticket);
    __syncthreads();
    sl = chain_decode(*s_ticket, cc.chunk_rows * TILES, cc.n_chunks);
    const int y = sl.chunk * cc.chunk_rows + sl.pos / TILES;
    tile = sl.pos % TILES;
    mid = -1;
    if (y < cc.rows) {
        rref.j = y / map.n_blocks; rref.blk = y - rref.j * map.n_blocks;
        mid = map.mod_id_of(rref.j, ntt_dig(map, rref.blk));
    }
    return mid >= 0;
}
FHE_D void chain_signal(const ChainCtl& cc, int chunk) {
    __syncthreads();                                   // the CTA's tile is stored ...
    if (threadIdx.x == 0) chain_red_release(cc.done + chunk);      // ... and published
}
FHE_D void chain_wait(const ChainCtl& cc, int chunk, unsigned need) {
    if (threadIdx.x == 0) {
        unsigned spins = 0;
        while (chain_ld_acquire(cc.done + chunk) < need) {
            chain_pause();
            if (++spins > FHE_CHAIN_SPIN_LIMIT) { *cc.err = 2u; break; }
        }
    }
    __syncthreads();
}

template <int LOG_R, class LoadOp, class StoreOp>
__global__ void __launch_bounds__(256, FHE_CHAIN_MINB) ntt_fwd_chained(DevTables T, RowMap map, LoadOp ld, StoreRawCG sp,
                                                                      LoadRawCG lp, StoreOp st, ChainCtl cc) {
    constexpr int TILES = 1 << (LOG_R - 4);
    FHE_SHARED __align__(16) double sm[16 * FHE_ROW_STRIDE];
    FHE_SHARED unsigned s_ticket;
    ChainSlot sl; RowRef rref; int mid, tile;
    const bool valid = chain_begin<LOG_R>(map, cc, sl, rref, mid, tile, &s_ticket);
    if (!sl.second) {
        if (valid) fwd_pass_a_body<LOG_R>(T, map, rref, mid, tile, ld, sp, sm);
        chain_signal(cc, sl.chunk);                    // padding and skipped rows count as well
    } else {
        chain_wait(cc, sl.chunk, (unsigned)(cc.chunk_rows * TILES));
        if (valid) fwd_pass_b_body<LOG_R + 8>(T, map, rref, mid, tile, lp, st, sm);
    }
}

template <int LOG_R, class LoadOp, class StoreOp>
__global__ void __launch_bounds__(256, FHE_CHAIN_MINB) ntt_inv_chained(DevTables T, RowMap map, LoadOp ld, StoreRawCG sp,
                                                                      LoadRawCG lp, StoreOp st, const ConstF* scale,
                                                                      ChainCtl cc) {
    constexpr int TILES = 1 << (LOG_R - 4);
    FHE_SHARED __align__(16) double sm[16 * FHE_ROW_STRIDE];
    FHE_SHARED unsigned s_ticket;
    ChainSlot sl; RowRef rref; int mid, tile;
    const bool valid = chain_begin<LOG_R>(map, cc, sl, rref, mid, tile, &s_ticket);
    if (!sl.second) {
        if (valid) inv_pass_b_body<LOG_R + 8>(T, map, rref, mid, tile, ld, sp, sm);
        chain_signal(cc, sl.chunk);
    } else {
        chain_wait(cc, sl.chunk, (unsigned)(cc.chunk_rows * TILES));
        if (valid) inv_pass_a_body<LOG_R>(T, map, rref, mid, tile, lp, st, scale, sm);
    }
}

// ------------------------------------------------------------------ host side
struct ChainHost {
    unsigned* ctr = nullptr;       // [1 + FHE_CHAIN_MAX_CHUNKS + 1]: ticket, done[], error flag
    int enabled = 0;
    int chunk_rows = FHE_CHAIN_ROWS;
};

inline bool chain_fill(ChainCtl& cc, const ChainHost& ch, int rows, cudaStream_t s) {
    if (!ch.enabled || !ch.ctr || rows < 1) return false;
    cc.rows = rows; cc.chunk_rows = ch.chunk_rows < rows ? ch.chunk_rows : rows;
    cc.n_chunks = (rows + cc.chunk_rows - 1) / cc.chunk_rows;
    if (cc.n_chunks > FHE_CHAIN_MAX_CHUNKS) return false;
    cc.ticket = ch.ctr; cc.done = ch.ctr + 1; cc.err = ch.ctr + 1 + FHE_CHAIN_MAX_CHUNKS;
    cudaMemsetAsync(ch.ctr, 0, sizeof(unsigned) * (1 + cc.n_chunks), s);
    return true;
}

template <int LOG_R, class LoadOp, class StoreOp>
inline bool ntt_forward_chained_r(const DevTables& T, const ChainHost& ch, const RowMap& map, int rows, LoadOp ld, u64* work,
                                  long long work_stride, StoreOp st, cudaStream_t s) {
    ChainCtl cc;
    if (!chain_fill(cc, ch, rows, s)) return false;
    StoreRawCG sp; sp.dst = work; sp.poly_stride = work_stride;
    LoadRawCG lp; lp.src = work; lp.poly_stride = work_stride;
    const unsigned ctas = 2u * cc.n_chunks * cc.chunk_rows * (1u << (LOG_R - 4));
    fhe_launch(ntt_fwd_chained<LOG_R, LoadOp, StoreOp>, dim3(ctas), dim3(256), 0, s, T, map, ld, sp, lp, st, cc);
    return true;
}
template <int LOG_R, class LoadOp, class StoreOp>
inline bool ntt_inverse_chained_r(const DevTables& T, const ChainHost& ch, const RowMap& map, int rows, LoadOp ld, u64* work,
                                  long long work_stride, StoreOp st, const ConstF* scale, cudaStream_t s) {
    ChainCtl cc;
    if (!chain_fill(cc, ch, rows, s)) return false;
    StoreRawCG sp; sp.dst = work; sp.poly_stride = work_stride;
    LoadRawCG lp; lp.src = work; lp.poly_stride = work_stride;
    const unsigned ctas = 2u * cc.n_chunks * cc.chunk_rows * (1u << (LOG_R - 4));
    fhe_launch(ntt_inv_chained<LOG_R, LoadOp, StoreOp>, dim3(ctas), dim3(256), 0, s, T, map, ld, sp, lp, st, scale, cc);
    return true;
}

template <class LoadOp, class StoreOp>
inline bool ntt_forward_chained(const DevTables& T, const ChainHost& ch, const RowMap& map_in, int rows, LoadOp ld, u64* work,
                                long long work_stride, StoreOp st, cudaStream_t s) {
    RowMap map = map_in; map.n_blocks = rows / map.rows_per_poly;
    switch (T.log_n - 8) {
        case 4: return ntt_forward_chained_r<4>(T, ch, map, rows, ld, work, work_stride, st, s);
        case 5: return ntt_forward_chained_r<5>(T, ch, map, rows, ld, work, work_stride, st, s);
        case 6: return ntt_forward_chained_r<6>(T, ch, map, rows, ld, work, work_stride, st, s);
        case 7: return ntt_forward_chained_r<7>(T, ch, map, rows, ld, work, work_stride, st, s);
        default: return ntt_forward_chained_r<8>(T, ch, map, rows, ld, work, work_stride, st, s);
    }
}
template <class LoadOp, class StoreOp>
inline bool ntt_inverse_chained(const DevTables& T, const ChainHost& ch, const RowMap& map_in, int rows, LoadOp ld, u64* work,
                                long long work_stride, StoreOp st, const ConstF* scale, cudaStream_t s) {
    RowMap map = map_in; map.n_blocks = rows / map.rows_per_poly;
    switch (T.log_n - 8) {
        case 4: return ntt_inverse_chained_r<4>(T, ch, map, rows, ld, work, work_stride, st, scale, s);
        case 5: return ntt_inverse_chained_r<5>(T, ch, map, rows, ld, work, work_stride, st, scale, s);
        case 6: return ntt_inverse_chained_r<6>(T, ch, map, rows, ld, work, work_stride, st, scale, s);
        case 7: return ntt_inverse_chained_r<7>(T, ch, map, rows, ld, work, work_stride, st, scale, s);
        default: return ntt_inverse_chained_r<8>(T, ch, map, rows, ld, work, work_stride, st, scale, s);
    }
}
