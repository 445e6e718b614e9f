// capi.cu -- context (device tables) and the extern "C" entry points declared in
// include/aesfhe_b200.h.  Host code here only builds small integer tables and sequences
// kernel launches on the caller's stream.
#include "kernels.cuh"
#include "../../include/aesfhe_b200.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <string>
#include <unordered_map>
#include <vector>
#include <cmath>

#ifndef FHE_BCONV_DEFAULT_VARIANT
#define FHE_BCONV_DEFAULT_VARIANT 2
#endif

typedef unsigned __int128 u128;

static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches(0);
static std::atomic<uint64_t> g_ntt_rows(0);          // length-N transforms (forward + inverse) launched since load

namespace {

// ------------------------------------------------------------------ host modular helpers
inline u64 h_mul(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }
inline u64 h_pow(u64 b, u64 e, u64 q) {
    u64 r = 1; b %= q;
    while (e) { if (e & 1) r = h_mul(r, b, q); b = h_mul(b, b, q); e >>= 1; }
    return r;
}
inline u64 h_inv(u64 a, u64 q) { return h_pow(a, q - 2, q); }
inline ConstF h_shoup(u64 w, u64 q) {
    ConstF s; s.w = (double)w; s.wq = (double)((long double)w / (long double)q); return s;
}
inline Modulus h_modulus(u64 q) {
    Modulus m; m.q = q; m.qd = (double)q; m.qinv = (double)(1.0L / (long double)q);
    u128 hi = ((u128)1 << 64) / q, rem = ((u128)1 << 64) % q;
    u128 lo = (rem << 64) / q;
    u128 mu = (hi << 64) + lo;
    m.mu_hi = (u64)(mu >> 64); m.mu_lo = (u64)mu;
    return m;
}
inline u32 h_brev(u32 x, int bits) {
    u32 r = 0;
    for (int i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

template <class T>
T* to_device(const std::vector<T>& v) {
    T* d = nullptr;
    if (cudaMalloc((void**)&d, sizeof(T) * (v.size() ? v.size() : 1)) != cudaSuccess) return nullptr;
    cudaMemcpy(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice);
    return d;
}

}  // namespace

struct fhe_ctx {
    int log_n, n, n_q, n_p, alpha, tot, device;
    std::vector<u64> q;                       // all moduli
    DevTables T;
    // per active-limb-count tables, index nq (1..n_q)
    std::vector<BConvTable*> modup_tables;    // device array of beta(nq) tables
    std::vector<int> modup_beta;
    std::vector<ConstF*> modup_scale;     // [nq][2] iNTT final constants incl. (Q_j/q_i)^-1
    std::vector<BConvTable*> moddown_table;   // one table
    std::vector<ConstF*> rescale_c;       // [nq-1] q_{nq-1}^-1 mod q_i
    // merged ModDown + rescale (divide by P * q_{nq-1}): sources = {q_{nq-1}} U P
    std::vector<BConvTable*> mdrs_table;  // one table, nt = nq-1
    std::vector<ConstF*> mdrs_scale;      // [K+1][2] iNTT final constants incl. (P q_last / m_k)^-1
    std::vector<ConstF*> mdrs_inv;        // [nq-1] (P q_last)^-1 mod q_i
    ConstF* p_mod_q = nullptr;            // [n_q]  P mod q_i
    ConstF* moddown_scale = nullptr;      // [n_p][2]
    ConstF* pinv = nullptr;               // [n_q]  P^-1 mod q_i
    u64 q0inv_mod_q1 = 0;
    std::vector<void*> owned;
    // host mirrors of the base-conversion tables, keyed by their device array: k_bconv_param takes its table by value
    std::unordered_map<const BConvTable*, std::vector<BConvTable>> bconv_host;
    std::unordered_map<const BConvTable*, std::vector<std::vector<unsigned char>>> bconv_params;   // BConvParam<ns> per table, built on first use
    // scratch arena
    u64* scratch = nullptr;
    size_t scratch_words = 0;
    // single-launch NTT (ntt_fused.cuh): L2-resident scratch limbs + group counters
    FusedHost fz;
    bool fz_calibrate = false;      // calibration still to be done when the fused path is first enabled

    u64 id_mod(int t, int nq) const { return q[t < nq ? t : n_q + (t - nq)]; }
    int id_of(int t, int nq) const { return t < nq ? t : n_q + (t - nq); }
};

namespace {

u64* arena(fhe_ctx* c, size_t words) {
    if (words > c->scratch_words) {
        if (c->scratch) { cudaDeviceSynchronize(); cudaFree(c->scratch); }
        c->scratch = nullptr;
        if (cudaMalloc((void**)&c->scratch, words * sizeof(u64)) != cudaSuccess) { c->scratch_words = 0; return nullptr; }
        c->scratch_words = words;
    }
    return c->scratch;
}

int fail(const char* what) {
    g_err = what;
    return -1;
}
int check(const char* where) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_err = std::string(where) + ": " + cudaGetErrorString(e);
        return -2;
    }
    return 0;
}

template <typename... KArgs, typename... Args>
inline void launch(void (*k)(KArgs...), dim3 g, dim3 b, cudaStream_t s, Args&&... args) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    fhe_launch(k, g, b, 0, s, std::forward<Args>(args)...);
}

template <class LoadOp, class StoreOp>
inline void ntt_fwd(fhe_ctx* c, const RowMap& m, int rows, LoadOp ld, u64* work, long long ws, StoreOp st, cudaStream_t s) {
    g_ntt_rows.fetch_add((uint64_t)rows, std::memory_order_relaxed);
    g_launches.fetch_add(ntt_forward_auto(c->T, c->fz, m, rows, ld, work, ws, st, s), std::memory_order_relaxed);
}
template <class LoadOp, class StoreOp>
inline void ntt_inv(fhe_ctx* c, const RowMap& m, int rows, LoadOp ld, u64* work, long long ws, StoreOp st,
                    const ConstF* scale, cudaStream_t s) {
    g_ntt_rows.fetch_add((uint64_t)rows, std::memory_order_relaxed);
    g_launches.fetch_add(ntt_inverse_auto(c->T, c->fz, m, rows, ld, work, ws, st, scale, s), std::memory_order_relaxed);
}

RowMap make_map(const fhe_ctx* c, int rows_per_poly, int j0, int nq, int skip_alpha = 0) {
    RowMap m;
    m.rows_per_poly = rows_per_poly; m.j0 = j0; m.nq = nq; m.p_base = c->n_q; m.skip_alpha = skip_alpha;
    m.digits = skip_alpha > 0 ? (nq + skip_alpha - 1) / skip_alpha : 1;
    m.log_n = c->log_n;
    m.n_blocks = 1;
    return m;
}

bool bad_shape(const fhe_ctx* c, int nq, int np) {
    return !c || nq < 1 || nq > c->n_q || (np != 0 && np != c->n_p);
}

// product of q[ids] except index `skip` (position in ids), reduced mod m
u64 prod_except(const fhe_ctx* c, const std::vector<int>& ids, int skip, u64 m) {
    u64 r = 1 % m;
    for (size_t i = 0; i < ids.size(); ++i)
        if ((int)i != skip) r = h_mul(r, c->q[ids[i]] % m, m);
    return r;
}

void build_level_tables(fhe_ctx* c) {
    const int n_q = c->n_q, K = c->n_p, alpha = c->alpha;
    c->modup_tables.assign(n_q + 1, nullptr);
    c->modup_beta.assign(n_q + 1, 0);
    c->modup_scale.assign(n_q + 1, nullptr);
    c->moddown_table.assign(n_q + 1, nullptr);
    c->rescale_c.assign(n_q + 1, nullptr);
    c->mdrs_table.assign(n_q + 1, nullptr);
    c->mdrs_scale.assign(n_q + 1, nullptr);
    c->mdrs_inv.assign(n_q + 1, nullptr);

    std::vector<u64> ninv(c->tot), w1ninv(c->tot);
    {
        // psi^-bitrev(1) = psi^-(N/2)
        std::vector<ConstF> fin(2 * (size_t)c->tot);
        cudaMemcpy(fin.data(), c->T.inv_final, sizeof(ConstF) * fin.size(), cudaMemcpyDeviceToHost);
        for (int l = 0; l < c->tot; ++l) { ninv[l] = (u64)fin[2 * l].w; w1ninv[l] = (u64)fin[2 * l + 1].w; }
    }
    std::vector<int> pids(K);
    for (int k = 0; k < K; ++k) pids[k] = n_q + k;

    for (int nq = 1; nq <= n_q; ++nq) {
        const int beta = (nq + alpha - 1) / alpha, ne = nq + K;
        c->modup_beta[nq] = beta;
        std::vector<BConvTable> tabs(beta);
        std::vector<ConstF> scale(2 * (size_t)nq);
        for (int j = 0; j < beta; ++j) {
            BConvTable& tb = tabs[j];
            std::memset(&tb, 0, sizeof(tb));
            const int lo = j * alpha, hi = std::min(lo + alpha, nq);
            std::vector<int> src;
            for (int i = lo; i < hi; ++i) src.push_back(i);
            tb.ns = (int)src.size();
            for (int k = 0; k < tb.ns; ++k) {
                tb.src_mod[k] = src[k]; tb.src_slot[k] = src[k];
                const u64 qk = c->q[src[k]];
                const u64 inv = h_inv(prod_except(c, src, k, qk), qk);
                scale[2 * (size_t)src[k]] = h_shoup(h_mul(ninv[src[k]], inv, qk), qk);
                scale[2 * (size_t)src[k] + 1] = h_shoup(h_mul(w1ninv[src[k]], inv, qk), qk);
            }
            int nt = 0;
            for (int t = 0; t < ne; ++t) {
                if (t >= lo && t < hi) continue;
                const int id = c->id_of(t, nq);
                const u64 m = c->q[id];
                tb.dst_mod[nt] = id; tb.dst_slot[nt] = t;
                for (int k = 0; k < tb.ns; ++k) tb.f[nt][k] = h_shoup(prod_except(c, src, k, m), m);
                ++nt;
            }
            tb.nt = nt;
        }
        c->modup_tables[nq] = to_device(tabs);
        c->bconv_host[c->modup_tables[nq]] = tabs;
        c->modup_scale[nq] = to_device(scale);
        c->owned.push_back(c->modup_tables[nq]); c->owned.push_back(c->modup_scale[nq]);

        if (K > 0) {
            std::vector<BConvTable> td(1);
            BConvTable& tb = td[0];
            std::memset(&tb, 0, sizeof(tb));
            tb.ns = K; tb.nt = nq;
            for (int k = 0; k < K; ++k) { tb.src_mod[k] = pids[k]; tb.src_slot[k] = k; }
            for (int i = 0; i < nq; ++i) {
                const u64 m = c->q[i];
                tb.dst_mod[i] = i; tb.dst_slot[i] = i;
                for (int k = 0; k < K; ++k) tb.f[i][k] = h_shoup(prod_except(c, pids, k, m), m);
            }
            c->moddown_table[nq] = to_device(td);
            c->bconv_host[c->moddown_table[nq]] = td;
            c->owned.push_back(c->moddown_table[nq]);
        }
        if (nq >= 2 && K > 0) {
            std::vector<int> sids;
            sids.push_back(nq - 1);
            for (int k = 0; k < K; ++k) sids.push_back(pids[k]);
            const int ns = K + 1;
            std::vector<BConvTable> td(1);
            BConvTable& tb = td[0];
            std::memset(&tb, 0, sizeof(tb));
            tb.ns = ns; tb.nt = nq - 1;
            std::vector<ConstF> sc(2 * (size_t)ns), inv(nq - 1);
            for (int k = 0; k < ns; ++k) {
                tb.src_mod[k] = sids[k]; tb.src_slot[k] = k;
                const u64 mk = c->q[sids[k]];
                const u64 iv = h_inv(prod_except(c, sids, k, mk), mk);
                sc[2 * (size_t)k] = h_shoup(h_mul(ninv[sids[k]], iv, mk), mk);
                sc[2 * (size_t)k + 1] = h_shoup(h_mul(w1ninv[sids[k]], iv, mk), mk);
            }
            for (int i = 0; i < nq - 1; ++i) {
                const u64 m = c->q[i];
                tb.dst_mod[i] = i; tb.dst_slot[i] = i;
                for (int k = 0; k < ns; ++k) tb.f[i][k] = h_shoup(prod_except(c, sids, k, m), m);
                inv[i] = h_shoup(h_inv(prod_except(c, sids, -1, m), m), m);
            }
            c->mdrs_table[nq] = to_device(td);
            c->bconv_host[c->mdrs_table[nq]] = td;
            c->mdrs_scale[nq] = to_device(sc);
            c->mdrs_inv[nq] = to_device(inv);
            c->owned.push_back(c->mdrs_table[nq]); c->owned.push_back(c->mdrs_scale[nq]); c->owned.push_back(c->mdrs_inv[nq]);
        }
        if (nq >= 2) {
            std::vector<ConstF> rc(nq - 1);
            const u64 ql = c->q[nq - 1];
            for (int i = 0; i < nq - 1; ++i) rc[i] = h_shoup(h_inv(ql % c->q[i], c->q[i]), c->q[i]);
            c->rescale_c[nq] = to_device(rc);
            c->owned.push_back(c->rescale_c[nq]);
        }
    }
    if (K > 0) {
        // iNTT final constants for the special limbs, replicated for up to 4 polys
        std::vector<ConstF> ms(2 * (size_t)K);
        for (int k = 0; k < K; ++k) {
            const u64 pk = c->q[pids[k]];
            const u64 inv = h_inv(prod_except(c, pids, k, pk), pk);
            ms[2 * (size_t)k] = h_shoup(h_mul(ninv[pids[k]], inv, pk), pk);
            ms[2 * (size_t)k + 1] = h_shoup(h_mul(w1ninv[pids[k]], inv, pk), pk);
        }
        c->moddown_scale = to_device(ms);
        c->owned.push_back(c->moddown_scale);
        std::vector<ConstF> pv(n_q);
        for (int i = 0; i < n_q; ++i) pv[i] = h_shoup(h_inv(prod_except(c, pids, -1, c->q[i]), c->q[i]), c->q[i]);
        c->pinv = to_device(pv);
        c->owned.push_back(c->pinv);
        std::vector<ConstF> pm(n_q);
        for (int i = 0; i < n_q; ++i) pm[i] = h_shoup(prod_except(c, pids, -1, c->q[i]), c->q[i]);
        c->p_mod_q = to_device(pm);
        c->owned.push_back(c->p_mod_q);
    }
    if (n_q >= 2) c->q0inv_mod_q1 = h_inv(c->q[0] % c->q[1], c->q[1]);
}

// the table of one launch as a kernel parameter (constants split at 2^23 on the host: the same h = rint(c 2^-23),
// l = c - h 2^23, s = h + l as split23 on the device)
template <int NS>
BConvParam<NS> bconv_param(const fhe_ctx* c, const BConvTable& tb) {
    BConvParam<NS> p;
    std::memset(&p, 0, sizeof(p));
    p.nt = tb.nt;
    for (int k = 0; k < NS; ++k) { p.src_slot[k] = tb.src_slot[k]; p.src_q[k] = c->q[tb.src_mod[k]]; }
    for (int t = 0; t < tb.nt; ++t) {
        p.dst_slot[t] = tb.dst_slot[t];
        const Modulus M = h_modulus(c->q[tb.dst_mod[t]]);
        p.q[t] = M.qd; p.qi[t] = M.qinv;
        for (int k = 0; k < NS; ++k) {
            const double v = tb.f[t][k].w;
            Split3 sp;
            sp.h = std::nearbyint(v * FHE_INV_TWO23);
            sp.l = v - sp.h * FHE_TWO23;
            sp.s = sp.h + sp.l;
            p.f[t][k] = sp;
        }
    }
    for (int t = tb.nt; t < FHE_MAX_DST; ++t) { p.q[t] = 3.0; p.qi[t] = 1.0 / 3.0; }     // never stored
    return p;
}
template <int NS>
void launch_bconv_ns(fhe_ctx* c, cudaStream_t s, int groups, const BConvTable* tabs, int n_tabs, u64* dst,
                     long long dst_stride, const u64* src, long long src_stride, int g_first, int g_step) {
    // variant 0: k_bconv (mulmod per term, table in shared memory); 2 (default): k_bconv_param (three-FMA dot product, table
    // as kernel parameter).  env FHE_BCONV_VARIANT
    static const int variant = getenv("FHE_BCONV_VARIANT") ? atoi(getenv("FHE_BCONV_VARIANT")) : FHE_BCONV_DEFAULT_VARIANT;
    if constexpr (NS <= 12) if (variant == 2 && g_step == n_tabs) {
        auto it = c->bconv_host.find(tabs);
        if (it != c->bconv_host.end()) {
            auto& vec = c->bconv_params[tabs];
            if (vec.empty()) vec.resize(it->second.size());
            std::vector<unsigned char>& blob = vec[g_first % n_tabs];
            if (blob.empty()) {                            // the table at index g_first has exactly NS sources (launch_bconv)
                const BConvParam<NS> p = bconv_param<NS>(c, it->second[g_first % n_tabs]);
                blob.resize(sizeof(p));
                std::memcpy(blob.data(), &p, sizeof(p));
            }
            constexpr int C = NS <= 7 ? 2 : 1;
            launch(k_bconv_param<NS, C>, dim3(c->n / (256 * C), groups), dim3(256), s, c->log_n,
                   *reinterpret_cast<const BConvParam<NS>*>(blob.data()), n_tabs, dst, dst_stride, src, src_stride, g_first, g_step);
            return;
        }
    }
    launch(k_bconv<NS>, dim3(c->n / 512, groups), dim3(256), s, c->T, tabs, n_tabs, dst, dst_stride, src, src_stride, g_first,
               g_step);
}
// one launch over groups g = y * g_step + g_first (y < groups) whose tables all have exactly `ns` sources
void launch_bconv_exact(fhe_ctx* c, cudaStream_t s, int ns, int groups, const BConvTable* tabs, int n_tabs, u64* dst,
                        long long dst_stride, const u64* src, long long src_stride, int g_first, int g_step) {
#define FHE_BCONV_CASE(NS) case NS: launch_bconv_ns<NS>(c, s, groups, tabs, n_tabs, dst, dst_stride, src, src_stride, g_first, g_step); break;
    switch (ns) {
        FHE_BCONV_CASE(1) FHE_BCONV_CASE(2) FHE_BCONV_CASE(3) FHE_BCONV_CASE(4)
        FHE_BCONV_CASE(5) FHE_BCONV_CASE(6) FHE_BCONV_CASE(7) FHE_BCONV_CASE(8)
        FHE_BCONV_CASE(9) FHE_BCONV_CASE(10) FHE_BCONV_CASE(11) FHE_BCONV_CASE(12)
        FHE_BCONV_CASE(13) FHE_BCONV_CASE(14) FHE_BCONV_CASE(15) FHE_BCONV_CASE(16)
        default: break;                                   // FHE_MAX_SRC = 16, checked at context creation
    }
#undef FHE_BCONV_CASE
}
// `src_counts[j]`: sources of table j (a ModUp launches its digits one by one, each with its exact digit size)
void launch_bconv(fhe_ctx* c, cudaStream_t s, const int* src_counts, int groups, const BConvTable* tabs, int n_tabs, u64* dst,
                  long long dst_stride, const u64* src, long long src_stride) {
    for (int j = 0; j < n_tabs; ++j)
        launch_bconv_exact(c, s, src_counts[j], groups / n_tabs, tabs, n_tabs, dst, dst_stride, src, src_stride, j, n_tabs);
}

// ab: `d` / `lift` are the operand ciphertexts a / b of a fused multiply (k_ks_inner, AB)
// lift_polys / accum / lift_only: see k_ks_inner (lift_only launches it with beta = 0)
void launch_ks_inner(fhe_ctx* c, cudaStream_t s, int nq, int batch, u64* acc, const u64* ext, const u64* d,
                     const u64* ksk, const u64* lift, const ConstF* lift_c, bool ab = false, int d_nq = 0,
                     int lift_nq = 0, int lift_polys = 2, int accum = 0, bool lift_only = false) {
    if (d_nq == 0) d_nq = nq;
    if (lift_nq == 0) lift_nq = nq;
    const int beta = lift_only ? 0 : c->modup_beta[nq];
    dim3 grid(c->n / 256, nq + c->n_p), block(256);
#define FHE_KS_ARGS grid, block, s, c->T, nq, c->alpha, batch, acc, ext, d, ksk, lift, lift_c, d_nq, lift_nq, \
                    lift_polys, accum
#define FHE_KS_CASE(BE, UN) case BE: if (ab) launch(k_ks_inner<BE, UN, true>, FHE_KS_ARGS); \
                                     else launch(k_ks_inner<BE, UN, false>, FHE_KS_ARGS); break;
    switch (beta) {
        case 0: launch(k_ks_inner<0, 2, false>, FHE_KS_ARGS); break;
        FHE_KS_CASE(1, 2) FHE_KS_CASE(2, 2) FHE_KS_CASE(3, 2) FHE_KS_CASE(4, 2)
        FHE_KS_CASE(5, 1) FHE_KS_CASE(6, 1) FHE_KS_CASE(7, 1)
        default: if (ab) launch(k_ks_inner<8, 1, true>, FHE_KS_ARGS); else launch(k_ks_inner<8, 1, false>, FHE_KS_ARGS); break;
    }
#undef FHE_KS_CASE
#undef FHE_KS_ARGS
}
#ifndef FHE_EMU
// The warp scheduler favours the CTAs that have been resident longest, so the groups of the
// persistent fused-NTT grid run at different speeds depending on the residency slot their CTAs
// occupy.  Measure it once (12 rows per group, equal split, wall time per group) and let
// fused_fill_ctl hand out row ranges proportional to the measured speeds.
void fused_calibrate(fhe_ctx* c) {
    FusedHost& fz = c->fz;
    const int groups = fused_full_groups(c->log_n, fz);
    if (groups < 2 || groups > FHE_FUSED_MAX_WGROUPS) return;
    const int per = 32, rows = groups * per;
    u64* buf = nullptr;
    if (cudaMalloc((void**)&buf, sizeof(u64) * (size_t)rows * c->n) != cudaSuccess) { cudaGetLastError(); return; }
    cudaMemset(buf, 0, sizeof(u64) * (size_t)rows * c->n);
    RowMap m = make_map(c, 1, 0, 1);
    LoadPlain ld; ld.src = buf; ld.poly_stride = c->n;
    StorePlain st; st.dst = buf; st.poly_stride = c->n;
    fz.calibrating = 1; fz.cal_groups = 0;
    std::vector<unsigned long long> t(groups);
    std::vector<int> start(groups + 1);
    // Fixed-point iteration: speeds interact (a late-slot group runs faster once the early-slot
    // groups have finished), so re-measure with the current split until the groups finish together.
    for (int it = 0; it < 6; ++it) {
        FusedCtl probe; fused_fill_ctl(probe, fz, groups, rows);
        for (int g = 0; g <= groups; ++g) start[g] = probe.weighted ? probe.start[g] : std::min(rows, g * per);
        bool ok = true;
        for (int rep = 0; rep < 2 && ok; ++rep) {
            cudaMemset(fz.gtime, 0, sizeof(unsigned long long) * FHE_FUSED_MAX_WGROUPS);
            ok = ntt_forward_auto(c->T, fz, m, rows, ld, buf, c->n, st, 0) == 1 && cudaDeviceSynchronize() == cudaSuccess;
        }
        if (ok) ok = cudaMemcpy(t.data(), fz.gtime, sizeof(unsigned long long) * groups, cudaMemcpyDeviceToHost) == cudaSuccess;
        for (int g = 0; g < groups && ok; ++g) if (t[g] == 0 || start[g + 1] <= start[g]) ok = false;
        if (!ok) { cudaGetLastError(); fz.cal_groups = 0; break; }
        // speed of a group = rows it was given / time it took (normalised); damped update
        std::vector<double> sp(groups);
        double tot = 0;
        for (int g = 0; g < groups; ++g) { sp[g] = (double)(start[g + 1] - start[g]) / (double)t[g]; tot += sp[g]; }
        for (int g = 0; g < groups; ++g) fz.weight[g] = fz.cal_groups ? 0.5 * fz.weight[g] + 0.5 * sp[g] / tot : sp[g] / tot;
        fz.cal_groups = groups;
    }
    fz.calibrating = 0;
    cudaFree(buf);
}
#endif

}  // namespace

// ModUp of [batch][nq][N] rows produced by `ld` (plain NTT-domain residues, or the product a1 * b1 of a fused multiply)
template <class LoadOp>
static int modup_from(fhe_ctx* c, cudaStream_t s, u64* ext, LoadOp ld, int nq, int batch) {
    const int n = c->n, ne = nq + c->n_p, beta = c->modup_beta[nq];
    u64* y = arena(c, (size_t)batch * nq * n);
    if (!y) return fail("fhe_modup: scratch allocation failed");
    {   // iNTT with the (Q_j/q_i)^-1 factor folded into the final stage
        RowMap m = make_map(c, nq, 0, nq);
        StorePlain st; st.dst = y; st.poly_stride = (long long)nq * n;
        ntt_inv(c, m, batch * nq, ld, y, (long long)nq * n, st, c->modup_scale[nq], s);
    }
    int digit_size[FHE_MAX_BETA];
    for (int j = 0; j < beta; ++j) digit_size[j] = std::min(c->alpha, nq - j * c->alpha);
    launch_bconv(c, s, digit_size, batch * beta, c->modup_tables[nq], beta, ext, (long long)ne * n, y, (long long)nq * n);
    {   // NTT of every converted row (a digit's own limbs are skipped)
        RowMap m = make_map(c, ne, 0, nq, c->alpha);
        LoadPlain l2; l2.src = (const u64*)ext; l2.poly_stride = (long long)ne * n;
        StorePlain st; st.dst = ext; st.poly_stride = (long long)ne * n;
        ntt_fwd(c, m, batch * beta * ne, l2, ext, (long long)ne * n, st, s);
        g_ntt_rows.fetch_sub((uint64_t)batch * nq, std::memory_order_relaxed);      // the skipped own-digit rows are not transforms
    }
    return check("fhe_modup");
}

// constant tiles above the default 48 KB dynamic shared-memory limit need an opt-in per kernel
#define FHE_LC_SMEM_LIMIT (160 * 1024)
template <class K> static void lc_allow_smem(K k, size_t smem) {
#ifndef FHE_EMU
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, FHE_LC_SMEM_LIMIT);
#else
    (void)k; (void)smem;
#endif
}

extern "C" {

const char* fhe_last_error(void) { return g_err.c_str(); }
uint64_t fhe_launch_count(void) { return g_launches.load(); }
uint64_t fhe_ntt_row_count(void) { return g_ntt_rows.load(); }

int fhe_ctx_create(fhe_ctx** out, int log_n, int n_q, int n_p, int alpha, const uint64_t* moduli,
                   const uint64_t* psi, int device) {
    if (!out || !moduli || !psi) return fail("fhe_ctx_create: null argument");
    if (log_n < 12 || log_n > 16) return fail("fhe_ctx_create: log_n must be in [12,16]");
    if (n_q < 1 || n_p < 0 || n_q + n_p > FHE_MAX_LIMBS || n_q + n_p > FHE_MAX_DST)
        return fail("fhe_ctx_create: too many limbs");
    if (alpha < 1 || alpha > FHE_MAX_SRC || n_p > FHE_MAX_SRC) return fail("fhe_ctx_create: digit size out of range");
    if (cudaSetDevice(device) != cudaSuccess) return fail("fhe_ctx_create: cudaSetDevice failed");
    fhe_ctx* c = new fhe_ctx;
    c->log_n = log_n; c->n = 1 << log_n; c->n_q = n_q; c->n_p = n_p; c->alpha = alpha; c->tot = n_q + n_p;
    c->device = device;
    const int n = c->n, tot = c->tot;
    c->q.assign(moduli, moduli + tot);
    std::vector<Modulus> mods(tot);
    std::vector<double> twf((size_t)tot * n), twi((size_t)tot * n);
    std::vector<ConstF> fin(2 * (size_t)tot);
    for (int l = 0; l < tot; ++l) {
        const u64 q = (u64)moduli[l];
        if (q >= (1ull << 45) || (q - 1) % (2ull * n) != 0) { fhe_ctx_destroy(c); return fail("fhe_ctx_create: modulus must be < 2^45 and = 1 mod 2N"); }
        if (h_pow(psi[l], n, q) != q - 1) { fhe_ctx_destroy(c); return fail("fhe_ctx_create: psi is not a primitive 2N-th root"); }
        mods[l] = h_modulus(q);
        const u64 ipsi = h_inv(psi[l], q);
        std::vector<u64> pw(n), ipw(n);
        pw[0] = 1; ipw[0] = 1;
        for (int k = 1; k < n; ++k) { pw[k] = h_mul(pw[k - 1], psi[l], q); ipw[k] = h_mul(ipw[k - 1], ipsi, q); }
        for (int k = 0; k < n; ++k) {
            const u32 r = h_brev((u32)k, log_n);
            twf[(size_t)l * n + k] = (double)pw[r];
            twi[(size_t)l * n + k] = (double)ipw[r];
        }
        const u64 ninv = h_inv((u64)n % q, q);
        fin[2 * (size_t)l] = h_shoup(ninv, q);
        fin[2 * (size_t)l + 1] = h_shoup(h_mul(ninv, ipw[h_brev(1u, log_n)], q), q);
    }
    Modulus* d_mod = to_device(mods);
    double* d_twf = to_device(twf);
    double* d_twi = to_device(twi);
    ConstF* d_fin = to_device(fin);
    if (!d_mod || !d_twf || !d_twi || !d_fin) { fhe_ctx_destroy(c); return fail("fhe_ctx_create: device allocation failed"); }
    c->owned = {d_mod, d_twf, d_twi, d_fin};
    c->T.mod = d_mod; c->T.tw_fwd = d_twf; c->T.tw_inv = d_twi; c->T.inv_final = d_fin;
    c->T.log_n = log_n; c->T.n_q = n_q; c->T.n_p = n_p;
    build_level_tables(c);
    {   // single-launch NTT resources (ntt_fused.cuh); FHE_NTT_FUSED=0 keeps the two-pass kernels
#ifndef FHE_EMU
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
#else
        const int sms = 4, coop = 1;
#endif
        const int gs = std::max(1, n / FHE_FUSED_TILE);
        FusedHost& fz = c->fz;
        fz.sm_count = sms;
        fz.max_groups = std::max(1, sms * FHE_FUSED_MAX_OCC / gs);
        if (const char* fl = std::getenv("FHE_FUSED_FLAGS")) fz.flags = std::atoi(fl);
        // resources are always there; mode 1 (persistent groups) is opt-in (FHE_NTT_FUSED=1 or
        // fhe_set_ntt_fused): it is slower than the other two on small launches (profiles/r01_fused_ntt.md)
        const char* env = std::getenv("FHE_NTT_FUSED");
        fz.enabled = 0;
        if (coop && sms > 0) {
            if (cudaMalloc((void**)&fz.scratch, sizeof(u64) * 2 * (size_t)fz.max_groups * n) != cudaSuccess ||
                cudaMalloc((void**)&fz.ctr, sizeof(unsigned) * (32 * fz.max_groups + 1)) != cudaSuccess) {
                fhe_ctx_destroy(c); return fail("fhe_ctx_create: fused-NTT scratch allocation failed");
            }
            cudaMemset(fz.ctr, 0, sizeof(unsigned) * (32 * fz.max_groups + 1));
            c->owned.push_back(fz.scratch); c->owned.push_back(fz.ctr);
            if (cudaMalloc((void**)&fz.gtime, sizeof(unsigned long long) * FHE_FUSED_MAX_WGROUPS) == cudaSuccess)
                c->owned.push_back(fz.gtime);
#ifdef FHE_FUSED_PROFILE
            cudaMalloc((void**)&fz.prof, 64 + 32 * 2048); cudaMemset(fz.prof, 0, 64 + 32 * 2048); c->owned.push_back(fz.prof);
#endif
            if (cudaMalloc((void**)&fz.chain.ctr, sizeof(unsigned) * (FHE_CHAIN_MAX_CHUNKS + 2)) == cudaSuccess) {
                cudaMemset(fz.chain.ctr, 0, sizeof(unsigned) * (FHE_CHAIN_MAX_CHUNKS + 2));
                c->owned.push_back(fz.chain.ctr);
                if (const char* cr = std::getenv("FHE_CHAIN_ROWS")) fz.chain.chunk_rows = std::max(1, std::atoi(cr));
                // default: the chained single-launch transform (same wall time as two launches on B200, half
                // the DRAM traffic, a third fewer launches); FHE_NTT_FUSED=0 selects the two-pass kernels
                fz.chain.enabled = (!env || env[0] == '2') ? 1 : 0;
            }
#ifdef FHE_EMU
            fz.enabled = !(env && env[0] == '0');      // the host simulator always exercises the fused kernels
#else
            fz.enabled = (env && env[0] == '1') ? 1 : 0;
#endif
        }
    }
#ifndef FHE_EMU
    {
        const char* cal = std::getenv("FHE_FUSED_CALIBRATE");
        c->fz_calibrate = !(cal && cal[0] == '0');
        if (c->fz.enabled && c->fz.gtime && c->fz_calibrate) { fused_calibrate(c); c->fz_calibrate = false; }
    }
#endif
    *out = c;
    return check("fhe_ctx_create");
}

int fhe_set_ntt_fused(fhe_ctx* c, int enabled) {
    if (!c) return fail("fhe_set_ntt_fused: null context");
    if (enabled && !c->fz.scratch) return fail("fhe_set_ntt_fused: context was created without fused-NTT resources");
    c->fz.chain.enabled = (enabled == 2 && c->fz.chain.ctr) ? 1 : 0;
    if (enabled == 2) { c->fz.enabled = 0; return c->fz.chain.enabled ? 0 : fail("fhe_set_ntt_fused: chained NTT unavailable"); }
    c->fz.enabled = enabled ? 1 : 0;
#ifndef FHE_EMU
    if (enabled && c->fz_calibrate && c->fz.gtime) { c->fz_calibrate = false; fused_calibrate(c); c->fz.enabled = 1; }
#endif
    return 0;
}

#ifdef FHE_FUSED_PROFILE
extern "C" int fhe_fused_profile(fhe_ctx* c, unsigned long long* out8) {
    cudaDeviceSynchronize();
    cudaMemcpy(out8, c->fz.prof, 64 + 32 * 2048, cudaMemcpyDeviceToHost);
    cudaMemset(c->fz.prof, 0, 64 + 32 * 2048);
    return 0;
}
#endif

int fhe_ntt_fused_status(fhe_ctx* c) {
    if (!c) return fail("fhe_ntt_fused_status: null context");
    if (!c->fz.ctr) return 0;
    // the flags are written by kernels on the caller's (non-blocking) streams: wait for all of them first
    if (cudaDeviceSynchronize() != cudaSuccess) return check("fhe_ntt_fused_status");
    unsigned flag = 0;
    if (cudaMemcpy(&flag, c->fz.ctr + 32 * c->fz.max_groups, sizeof(unsigned), cudaMemcpyDeviceToHost) != cudaSuccess)
        return check("fhe_ntt_fused_status");
    if (flag) return fail("fused NTT: a group barrier timed out (CTAs of a group were not co-resident)");
    if (c->fz.chain.ctr) {
        if (cudaMemcpy(&flag, c->fz.chain.ctr + 1 + FHE_CHAIN_MAX_CHUNKS, sizeof(unsigned), cudaMemcpyDeviceToHost) != cudaSuccess)
            return check("fhe_ntt_fused_status");
        if (flag) return fail("chained NTT: a second-pass CTA timed out waiting for its chunk");
    }
    return 0;
}

void fhe_ctx_destroy(fhe_ctx* c) {
    if (!c) return;
    cudaDeviceSynchronize();
    for (void* p : c->owned) cudaFree(p);
    if (c->scratch) cudaFree(c->scratch);
    delete c;
}

int fhe_ntt_fwd(fhe_ctx* c, void* stream, uint64_t* data, int npoly, int nq, int np) {
    if (bad_shape(c, nq, np) || npoly < 1) return fail("fhe_ntt_fwd: bad shape");
    const int rpp = nq + np;
    RowMap m = make_map(c, rpp, 0, nq);
    const long long ps = (long long)rpp * c->n;
    LoadPlain ld; ld.src = (const u64*)data; ld.poly_stride = ps;
    StorePlain st; st.dst = (u64*)data; st.poly_stride = ps;
    ntt_fwd(c, m, npoly * rpp, ld, (u64*)data, ps, st, (cudaStream_t)stream);
    return check("fhe_ntt_fwd");
}

int fhe_ntt_inv(fhe_ctx* c, void* stream, uint64_t* data, int npoly, int nq, int np) {
    if (bad_shape(c, nq, np) || npoly < 1) return fail("fhe_ntt_inv: bad shape");
    const int rpp = nq + np;
    RowMap m = make_map(c, rpp, 0, nq);
    const long long ps = (long long)rpp * c->n;
    LoadPlain ld; ld.src = (const u64*)data; ld.poly_stride = ps;
    StorePlain st; st.dst = (u64*)data; st.poly_stride = ps;
    ntt_inv(c, m, npoly * rpp, ld, (u64*)data, ps, st, nullptr, (cudaStream_t)stream);
    return check("fhe_ntt_inv");
}

static int binary(int op, fhe_ctx* c, void* stream, uint64_t* out, const uint64_t* a, const uint64_t* b,
                  int npoly, int batch, int b_npoly, int b_batch, int nq, int np) {
    if (bad_shape(c, nq, np) || npoly < 1 || batch < 1 || (b_npoly != npoly && b_npoly != 1) ||
        (b_batch != batch && b_batch != 1))
        return fail("fhe_add/sub/mul: bad shape");
    const int rpp = nq + np;
    RowMap m = make_map(c, rpp, 0, nq);
    const long long blk = (long long)rpp * c->n;
    Strides so, sa, sb;
    so.poly = sa.poly = blk * batch; so.batch = sa.batch = blk;
    sb.batch = b_batch == 1 ? 0 : blk;
    sb.poly = b_npoly == 1 ? 0 : blk * b_batch;
    dim3 grid(c->n / 1024, npoly * batch * rpp), block(256);
    cudaStream_t s = (cudaStream_t)stream;
    if (op == 0) launch(k_binary<0>, grid, block, s, c->T, m, batch, (u64*)out, (const u64*)a, (const u64*)b, so, sa, sb);
    else if (op == 1) launch(k_binary<1>, grid, block, s, c->T, m, batch, (u64*)out, (const u64*)a, (const u64*)b, so, sa, sb);
    else launch(k_binary<2>, grid, block, s, c->T, m, batch, (u64*)out, (const u64*)a, (const u64*)b, so, sa, sb);
    return check("fhe_binary");
}
int fhe_add(fhe_ctx* c, void* s, uint64_t* o, const uint64_t* a, const uint64_t* b, int npoly, int batch, int b_npoly,
            int b_batch, int nq, int np) { return binary(0, c, s, o, a, b, npoly, batch, b_npoly, b_batch, nq, np); }
int fhe_sub(fhe_ctx* c, void* s, uint64_t* o, const uint64_t* a, const uint64_t* b, int npoly, int batch, int b_npoly,
            int b_batch, int nq, int np) { return binary(1, c, s, o, a, b, npoly, batch, b_npoly, b_batch, nq, np); }
int fhe_mul(fhe_ctx* c, void* s, uint64_t* o, const uint64_t* a, const uint64_t* b, int npoly, int batch, int b_npoly,
            int b_batch, int nq, int np) { return binary(2, c, s, o, a, b, npoly, batch, b_npoly, b_batch, nq, np); }

int fhe_neg(fhe_ctx* c, void* stream, uint64_t* out, const uint64_t* a, int npoly, int nq, int np) {
    if (bad_shape(c, nq, np) || npoly < 1) return fail("fhe_neg: bad shape");
    const int rpp = nq + np;
    launch(k_neg, dim3(c->n / 256, npoly * rpp), dim3(256), (cudaStream_t)stream, c->T, make_map(c, rpp, 0, nq),
           (u64*)out, (const u64*)a);
    return check("fhe_neg");
}

int fhe_tensor(fhe_ctx* c, void* stream, uint64_t* out, const uint64_t* a, const uint64_t* b, int nq, int batch) {
    if (bad_shape(c, nq, 0) || batch < 1) return fail("fhe_tensor: bad shape");
    launch(k_tensor, dim3(c->n / 256, batch * nq), dim3(256), (cudaStream_t)stream, c->T, nq, (u64*)out, (const u64*)a,
           (const u64*)b, (long long)batch * nq * c->n);
    return check("fhe_tensor");
}

static int constop(int add, fhe_ctx* c, void* stream, uint64_t* out, const uint64_t* a, const uint64_t* c1,
                   const uint64_t* c2, int npoly, int nq, int np) {
    if (bad_shape(c, nq, np) || npoly < 1 || !c1 || !c2) return fail("fhe_mul_const/add_const: bad shape");
    const int rpp = nq + np;
    LimbConsts k;
    for (int j = 0; j < rpp; ++j) { k.a[j] = c1[j]; k.b[j] = c2[j]; }
    dim3 grid(c->n / 256, npoly * rpp), block(256);
    RowMap m = make_map(c, rpp, 0, nq);
    if (add) launch(k_const<1>, grid, block, (cudaStream_t)stream, c->T, m, (u64*)out, (const u64*)a, k);
    else launch(k_const<0>, grid, block, (cudaStream_t)stream, c->T, m, (u64*)out, (const u64*)a, k);
    return check("fhe_const");
}
int fhe_mul_const(fhe_ctx* c, void* s, uint64_t* o, const uint64_t* a, const uint64_t* c1, const uint64_t* c2,
                  int npoly, int nq, int np) { return constop(0, c, s, o, a, c1, c2, npoly, nq, np); }
int fhe_add_const(fhe_ctx* c, void* s, uint64_t* o, const uint64_t* a, const uint64_t* c1, const uint64_t* c2,
                  int npoly, int nq, int np) { return constop(1, c, s, o, a, c1, c2, npoly, nq, np); }

int fhe_lincomb(fhe_ctx* c, void* stream, uint64_t* out, const uint64_t* const* in, const int* in_nq,
                const double* consts, const uint64_t* c0, int M, int T, int nq, int batch, const long long* in_poly_stride) {
    if (bad_shape(c, nq, 0) || batch < 1 || M < 1 || T < 1 || T > FHE_LC_MAX_T || !in || !in_nq || !consts)
        return fail("fhe_lincomb: bad shape");
    LinCombIn li;
    for (int t = 0; t < FHE_LC_MAX_T; ++t) {
        const int nt = t < T ? in_nq[t] : nq;
        if (t < T && nt < nq) return fail("fhe_lincomb: input has fewer limbs than the output");
        li.ptr[t] = t < T ? (const u64*)in[t] : nullptr;
        li.batch_stride[t] = (long long)nt * c->n;
        li.poly_stride[t] = (in_poly_stride && t < T) ? in_poly_stride[t] : (long long)batch * nt * c->n;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const ConstF* cf = (const ConstF*)consts;
    if (M <= 2) {                                            // too few columns for a tensor-core tile
        dim3 grid(c->n / 512, 2 * batch * nq), block(256);
        if (T <= 4) launch(k_lincomb_few<4>, grid, block, s, c->T, nq, batch, M, T, li, cf, (const u64*)c0, (u64*)out);
        else launch(k_lincomb_few<16>, grid, block, s, c->T, nq, batch, M, T, li, cf, (const u64*)c0, (u64*)out);
        return check("fhe_lincomb");
    }
    const int kt = (T + 3) / 4;
    const size_t smem = FHE_LCM_SMEM(M, kt);
    if (smem > FHE_LC_SMEM_LIMIT) return fail("fhe_lincomb: M * T too large for the shared-memory constant tile");
    dim3 grid(c->n / FHE_LCM_CHUNK, 2 * batch * nq), block(256);
    g_launches.fetch_add(1, std::memory_order_relaxed);
#define FHE_LC_CASE(KT) case KT: lc_allow_smem(k_lincomb_mma<KT>, smem); \
        fhe_launch(k_lincomb_mma<KT>, grid, block, smem, s, c->T, nq, batch, M, T, li, cf, (const u64*)c0, (u64*)out); break;
    switch (kt) { FHE_LC_CASE(1) FHE_LC_CASE(2) FHE_LC_CASE(3) default: FHE_LC_CASE(4) }
#undef FHE_LC_CASE
    return check("fhe_lincomb");
}

int fhe_mul_plain_sum(fhe_ctx* c, void* stream, uint64_t* out, const uint64_t* const* a, const int* a_nq,
                      const uint64_t* const* p, int T, int nq, int batch, int accumulate) {
    if (bad_shape(c, nq, 0) || batch < 1 || T < 1 || T > FHE_LC_MAX_T || !a || !a_nq || !p)
        return fail("fhe_mul_plain_sum: bad shape");
    PlainSumIn in;
    for (int t = 0; t < FHE_LC_MAX_T; ++t) {
        const int nt = t < T ? a_nq[t] : nq;
        if (t < T && nt < nq) return fail("fhe_mul_plain_sum: operand has fewer limbs than the output");
        in.a[t] = t < T ? (const u64*)a[t] : nullptr;
        in.p[t] = t < T ? (const u64*)p[t] : nullptr;
        in.a_batch_stride[t] = (long long)nt * c->n;
        in.a_poly_stride[t] = (long long)batch * nt * c->n;
    }
    dim3 grid(c->n / 256, nq), block(256);
    cudaStream_t s = (cudaStream_t)stream;
    if (T <= 4) launch(k_mul_plain_sum<4>, grid, block, s, c->T, nq, batch, T, in, (u64*)out, accumulate);
    else if (T <= 8) launch(k_mul_plain_sum<8>, grid, block, s, c->T, nq, batch, T, in, (u64*)out, accumulate);
    else launch(k_mul_plain_sum<16>, grid, block, s, c->T, nq, batch, T, in, (u64*)out, accumulate);
    return check("fhe_mul_plain_sum");
}

int fhe_mul_plain_multi(fhe_ctx* c, void* stream, uint64_t* out, const uint64_t* const* a, const int* a_nq,
                        const uint64_t* const* p, int T, int G, int nq, int batch) {
    if (bad_shape(c, nq, 0) || batch < 1 || T < 1 || T > FHE_LC_MAX_T || G < 1 || G > FHE_PM_MAX_G || !a || !a_nq || !p)
        return fail("fhe_mul_plain_multi: bad shape");
    PlainMultiIn in;
    for (int t = 0; t < FHE_LC_MAX_T; ++t) {
        const int nt = t < T ? a_nq[t] : nq;
        if (t < T && nt < nq) return fail("fhe_mul_plain_multi: operand has fewer limbs than the output");
        in.a[t] = t < T ? (const u64*)a[t] : nullptr;
        in.a_batch_stride[t] = (long long)nt * c->n;
        in.a_poly_stride[t] = (long long)batch * nt * c->n;
        for (int g = 0; g < FHE_PM_MAX_G; ++g) in.p[g][t] = (g < G && t < T) ? (const u64*)p[(size_t)g * T + t] : nullptr;
    }
    dim3 grid(c->n / 256, nq), block(256);
    cudaStream_t s = (cudaStream_t)stream;
    if (T <= 4) launch(k_mul_plain_multi<4>, grid, block, s, c->T, nq, batch, T, G, in, (u64*)out);
    else if (T <= 8) launch(k_mul_plain_multi<8>, grid, block, s, c->T, nq, batch, T, G, in, (u64*)out);
    else launch(k_mul_plain_multi<16>, grid, block, s, c->T, nq, batch, T, G, in, (u64*)out);
    return check("fhe_mul_plain_multi");
}

int fhe_tensor_acc(fhe_ctx* c, void* stream, uint64_t* acc, const uint64_t* const* a, const int* a_nq,
                   const int* a_batch, const uint64_t* b, int b_batch, int G, int nq, int batch, int accumulate,
                   const long long* a_poly_stride, long long acc_poly_stride, const uint64_t* init) {
    if (bad_shape(c, nq, 0) || batch < 1 || G < 1 || G > FHE_LC_MAX_T || !a || !a_nq || !a_batch || !b ||
        (b_batch != batch && b_batch != 1))
        return fail("fhe_tensor_acc: bad shape");
    TensorAccIn ti;
    for (int g = 0; g < FHE_LC_MAX_T; ++g) {
        const int ng = g < G ? a_nq[g] : nq;
        if (g < G && ng < nq) return fail("fhe_tensor_acc: operand has fewer limbs than the accumulator");
        const int ab = g < G ? a_batch[g] : batch;
        if (ab != batch && ab != 1) return fail("fhe_tensor_acc: operand batch must be 1 or the accumulator's");
        ti.a[g] = g < G ? (const u64*)a[g] : nullptr;
        ti.a_batch_stride[g] = ab == 1 ? 0 : (long long)ng * c->n;
        ti.a_poly_stride[g] = (a_poly_stride && g < G) ? a_poly_stride[g] : (long long)ab * ng * c->n;
    }
    if (acc_poly_stride != 0 && acc_poly_stride < (long long)batch * nq * c->n)
        return fail("fhe_tensor_acc: accumulator polynomial stride shorter than one polynomial");
    launch(k_tensor_acc, dim3(c->n / 256, batch * nq), dim3(256), (cudaStream_t)stream, c->T, nq, batch, G, ti,
           (const u64*)b, b_batch, (u64*)acc, accumulate, acc_poly_stride ? acc_poly_stride : (long long)batch * nq * c->n,
           (const u64*)init);
    return check("fhe_tensor_acc");
}

int fhe_rescale(fhe_ctx* c, void* stream, uint64_t* out, const uint64_t* in, int npoly, int nq) {
    if (bad_shape(c, nq, 0) || nq < 2 || npoly < 1) return fail("fhe_rescale: bad shape");
    cudaStream_t s = (cudaStream_t)stream;
    const int n = c->n;
    u64* last = arena(c, (size_t)npoly * n);
    if (!last) return fail("fhe_rescale: scratch allocation failed");
    // 1) dropped limb -> coefficient domain
    {
        RowMap m = make_map(c, 1, nq - 1, nq);
        LoadPlain ld; ld.src = (const u64*)in + (size_t)(nq - 1) * n; ld.poly_stride = (long long)nq * n;
        StorePlain st; st.dst = last; st.poly_stride = n;
        ntt_inv(c, m, npoly, ld, last, n, st, nullptr, s);
    }
    // 2) per remaining limb: NTT(centred remainder), out = (in - it) * q_last^-1
    {
        const int rpp = nq - 1;
        RowMap m = make_map(c, rpp, 0, nq);
        LoadCentered ld; ld.last = last; ld.q_last = c->q[nq - 1];
        StoreSubMul st; st.out = (u64*)out; st.out_poly_stride = (long long)rpp * n;
        st.in = (const u64*)in; st.in_poly_stride = (long long)nq * n; st.c = c->rescale_c[nq];
        ntt_fwd(c, m, npoly * rpp, ld, (u64*)out, (long long)rpp * n, st, s);
    }
    return check("fhe_rescale");
}

int fhe_mod_raise(fhe_ctx* c, void* stream, uint64_t* out, const uint64_t* in, int npoly, int nq_out) {
    if (bad_shape(c, nq_out, 0) || npoly < 1) return fail("fhe_mod_raise: bad shape");
    cudaStream_t s = (cudaStream_t)stream;
    const int n = c->n;
    u64* coef = arena(c, (size_t)npoly * n);
    if (!coef) return fail("fhe_mod_raise: scratch allocation failed");
    {   // level-0 rows -> coefficient domain
        RowMap m = make_map(c, 1, 0, 1);
        LoadPlain ld; ld.src = (const u64*)in; ld.poly_stride = n;
        StorePlain st; st.dst = coef; st.poly_stride = n;
        ntt_inv(c, m, npoly, ld, coef, n, st, nullptr, s);
    }
    {   // centred lift into every target modulus, fused into the forward transform's load
        RowMap m = make_map(c, nq_out, 0, nq_out);
        LoadCentered ld; ld.last = coef; ld.q_last = c->q[0];
        StorePlain st; st.dst = (u64*)out; st.poly_stride = (long long)nq_out * n;
        ntt_fwd(c, m, npoly * nq_out, ld, (u64*)out, (long long)nq_out * n, st, s);
    }
    return check("fhe_mod_raise");
}

int fhe_automorphism(fhe_ctx* c, void* stream, uint64_t* out, const uint64_t* in, uint64_t galois, int nrows) {
    if (!c || nrows < 1 || !(galois & 1)) return fail("fhe_automorphism: bad arguments");
    launch(k_automorphism, dim3(c->n / 1024, nrows), dim3(256), (cudaStream_t)stream, c->log_n, (u64*)out,
           (const u64*)in, (u64)galois);
    return check("fhe_automorphism");
}

int fhe_modup(fhe_ctx* c, void* stream, uint64_t* ext, const uint64_t* d, int nq, int batch) {
    if (bad_shape(c, nq, 0) || c->n_p == 0 || batch < 1) return fail("fhe_modup: bad shape");
    LoadPlain ld; ld.src = (const u64*)d; ld.poly_stride = (long long)nq * c->n;
    return modup_from(c, (cudaStream_t)stream, (u64*)ext, ld, nq, batch);
}

int fhe_ks_inner(fhe_ctx* c, void* stream, uint64_t* acc, const uint64_t* ext, const uint64_t* d,
                 const uint64_t* ksk, int nq, int batch) {
    if (bad_shape(c, nq, 0) || c->n_p == 0 || batch < 1 || c->modup_beta[nq] > FHE_MAX_BETA)
        return fail("fhe_ks_inner: bad shape");
    launch_ks_inner(c, (cudaStream_t)stream, nq, batch, (u64*)acc, (const u64*)ext, (const u64*)d, (const u64*)ksk,
                    nullptr, nullptr);
    return check("fhe_ks_inner");
}

int fhe_bsgs_inner(fhe_ctx* c, void* stream, uint64_t* out, const uint64_t* ext, const uint64_t* ct, int ct_nq,
                   const uint64_t* const* keys, const uint64_t* galois, const uint64_t* const* pts, int nb, int G, int nq,
                   int batch, int accumulate) {
    if (bad_shape(c, nq, 0) || c->n_p == 0 || batch < 1 || c->modup_beta[nq] > 4 || nb < 1 || nb > FHE_BSGS_MAX_BABY ||
        G < 1 || G > FHE_BSGS_MAX_G || ct_nq < nq || !out || !ext || !ct || !keys || !galois || !pts)
        return fail("fhe_bsgs_inner: bad arguments");
    BsgsIn in;
    for (int b = 0; b < FHE_BSGS_MAX_BABY; ++b) {
        in.key[b] = b < nb ? (const u64*)keys[b] : nullptr;
        in.galois[b] = b < nb ? (u64)galois[b] : 1;
        for (int g = 0; g < FHE_BSGS_MAX_G; ++g) in.pt[g][b] = (g < G && b < nb) ? (const u64*)pts[(size_t)g * nb + b] : nullptr;
    }
    const int beta = c->modup_beta[nq];
    cudaStream_t s = (cudaStream_t)stream;
    // variant 2 (default): accumulators in shared memory, 3 CTAs per SM: 4.8 ms at the CoeffToSlot shape against 5.5 ms for
    // variant 0 (accumulators in registers, 2 CTAs per SM); env FHE_BSGS_VARIANT, tools/bsgs_bench.py
    static const int variant = getenv("FHE_BSGS_VARIANT") ? atoi(getenv("FHE_BSGS_VARIANT")) : 2;
    constexpr int BB = 4;
    const dim3 grid((c->n / 256) * ((batch + BB - 1) / BB), nq + c->n_p), block(256);
#define FHE_BSGS_GO(BE, GN) do { \
        if (variant == 2) { const size_t smem = (size_t)GN * 2 * BB * 256 * sizeof(double); \
            lc_allow_smem(k_bsgs_inner<BE, GN, BB, true>, smem); g_launches.fetch_add(1, std::memory_order_relaxed); \
            fhe_launch(k_bsgs_inner<BE, GN, BB, true>, grid, block, smem, s, c->T, nq, c->alpha, batch, nb, in, (const u64*)ext, \
                       (const u64*)ct, ct_nq, (const ConstF*)c->p_mod_q, (u64*)out, accumulate); } \
        else launch(k_bsgs_inner<BE, GN, BB, false>, grid, block, s, c->T, nq, c->alpha, batch, nb, in, (const u64*)ext, \
                    (const u64*)ct, ct_nq, (const ConstF*)c->p_mod_q, (u64*)out, accumulate); } while (0)
#define FHE_BSGS_G(BE) switch (G) { case 1: FHE_BSGS_GO(BE, 1); break; case 2: FHE_BSGS_GO(BE, 2); break; \
                                    case 3: FHE_BSGS_GO(BE, 3); break; default: FHE_BSGS_GO(BE, 4); break; }
    switch (beta) {
        case 1: FHE_BSGS_G(1); break;
        case 2: FHE_BSGS_G(2); break;
        case 3: FHE_BSGS_G(3); break;
        default: FHE_BSGS_G(4); break;
    }
#undef FHE_BSGS_G
#undef FHE_BSGS_GO
    return check("fhe_bsgs_inner");
}

// ModDown by P * q_{nq-1} of an extended accumulator acc[npoly][nq + K][N] (used as scratch) -> out[npoly][nq-1][N]
static int moddown_rescale_tail(fhe_ctx* c, cudaStream_t s, u64* out, u64* acc, int nq, int npoly, const char* who) {
    const int n = c->n, K = c->n_p, ne = nq + K, no = nq - 1;
    u64* accs = acc + (size_t)no * n;                       // rows q_{nq-1}, p_0 .. p_{K-1} are contiguous
    {
        RowMap m = make_map(c, K + 1, no, nq);
        LoadPlain ld; ld.src = accs; ld.poly_stride = (long long)ne * n;
        StorePlain st; st.dst = accs; st.poly_stride = (long long)ne * n;
        ntt_inv(c, m, npoly * (K + 1), ld, accs, (long long)ne * n, st, c->mdrs_scale[nq], s);
    }
    { const int ns = K + 1; launch_bconv(c, s, &ns, npoly, c->mdrs_table[nq], 1, out, (long long)no * n, accs, (long long)ne * n); }
    {
        RowMap m = make_map(c, no, 0, nq);
        LoadPlain ld; ld.src = (const u64*)out; ld.poly_stride = (long long)no * n;
        StoreSubMul st; st.out = out; st.out_poly_stride = (long long)no * n;
        st.in = acc; st.in_poly_stride = (long long)ne * n; st.c = c->mdrs_inv[nq];
        ntt_fwd(c, m, npoly * no, ld, out, (long long)no * n, st, s);
    }
    return check(who);
}

// d3 != null: the 3-polynomial tensor product; else (a, b): the operands of a fused multiply
static int relin_rescale_impl(fhe_ctx* c, void* stream, uint64_t* out, const uint64_t* d3, const uint64_t* a, int a_nq,
                              const uint64_t* b, int b_nq, const uint64_t* rlk, int nq, int batch, const char* who) {
    if (bad_shape(c, nq, 0) || nq < 2 || c->n_p == 0 || batch < 1 || c->modup_beta[nq] > FHE_MAX_BETA)
        return fail((std::string(who) + ": bad shape").c_str());
    cudaStream_t s = (cudaStream_t)stream;
    const int n = c->n, K = c->n_p, ne = nq + K, beta = c->modup_beta[nq], no = nq - 1;
    const size_t words = ((size_t)batch * nq + (size_t)batch * beta * ne + 2 * (size_t)batch * ne) * n;
    u64* base = arena(c, words);
    if (!base) return fail((std::string(who) + ": scratch allocation failed").c_str());
    u64* ext = base + (size_t)batch * nq * n;
    u64* acc = ext + (size_t)batch * beta * ne * n;
    int rc;
    if (d3) {
        const u64* d2 = (const u64*)d3 + 2 * (size_t)batch * nq * n;
        if ((rc = fhe_modup(c, stream, (uint64_t*)ext, (const uint64_t*)d2, nq, batch))) return rc;
        // inner product with the key, plus P * (d0, d1) on the q-limbs
        launch_ks_inner(c, s, nq, batch, acc, (const u64*)ext, d2, (const u64*)rlk, (const u64*)d3, c->p_mod_q);
    } else {
        if (a_nq < nq || a_nq > c->n_q || b_nq < nq || b_nq > c->n_q)
            return fail((std::string(who) + ": operand has fewer limbs than the product").c_str());
        LoadMul ld;
        ld.a = (const u64*)a + (size_t)batch * a_nq * n; ld.a_stride = (long long)a_nq * n;
        ld.b = (const u64*)b + (size_t)batch * b_nq * n; ld.b_stride = (long long)b_nq * n;
        if ((rc = modup_from(c, s, ext, ld, nq, batch))) return rc;
        launch_ks_inner(c, s, nq, batch, acc, (const u64*)ext, (const u64*)a, (const u64*)rlk, (const u64*)b, c->p_mod_q,
                        true, a_nq, b_nq);
    }
    return moddown_rescale_tail(c, s, (u64*)out, acc, nq, 2 * batch, who);
}

int fhe_relin_rescale(fhe_ctx* c, void* stream, uint64_t* out, const uint64_t* d3, const uint64_t* rlk, int nq,
                      int batch) {
    if (!d3) return fail("fhe_relin_rescale: null operand");
    return relin_rescale_impl(c, stream, out, d3, nullptr, 0, nullptr, 0, rlk, nq, batch, "fhe_relin_rescale");
}

int fhe_mul_relin_rescale(fhe_ctx* c, void* stream, uint64_t* out, const uint64_t* a, int a_nq, const uint64_t* b,
                          int b_nq, const uint64_t* rlk, int nq, int batch) {
    if (!a || !b) return fail("fhe_mul_relin_rescale: null operand");
    return relin_rescale_impl(c, stream, out, nullptr, a, a_nq, b, b_nq, rlk, nq, batch, "fhe_mul_relin_rescale");
}

int fhe_mul_relin_rescale_ptrs(fhe_ctx* c, void* stream, uint64_t* out, const uint64_t* const* a0, const uint64_t* const* a1,
                               const uint64_t* const* b0, const uint64_t* const* b1, const uint64_t* rlk, int nq, int batch) {
    if (bad_shape(c, nq, 0) || nq < 2 || c->n_p == 0 || batch < 1 || batch > FHE_MAX_MULB || c->modup_beta[nq] > 4 ||
        !out || !a0 || !a1 || !b0 || !b1 || !rlk)
        return fail("fhe_mul_relin_rescale_ptrs: bad arguments (1 <= batch <= 128, at most four digits)");
    cudaStream_t s = (cudaStream_t)stream;
    const int n = c->n, K = c->n_p, ne = nq + K, beta = c->modup_beta[nq];
    const size_t words = ((size_t)batch * nq + (size_t)batch * beta * ne + 2 * (size_t)batch * ne) * n;
    u64* base = arena(c, words);
    if (!base) return fail("fhe_mul_relin_rescale_ptrs: scratch allocation failed");
    u64* ext = base + (size_t)batch * nq * n;
    u64* acc = ext + (size_t)batch * beta * ne * n;
    MulPtrs mp;
    LoadMulPtr ld;
    for (int b = 0; b < FHE_MAX_MULB; ++b) {
        const int i = b < batch ? b : batch - 1;
        if (!a0[i] || !a1[i] || !b0[i] || !b1[i]) return fail("fhe_mul_relin_rescale_ptrs: null operand pointer");
        mp.a0[b] = (const u64*)a0[i]; mp.a1[b] = (const u64*)a1[i]; mp.b0[b] = (const u64*)b0[i]; mp.b1[b] = (const u64*)b1[i];
        ld.a1[b] = mp.a1[b]; ld.b1[b] = mp.b1[b];
    }
    int rc;
    if ((rc = modup_from(c, s, ext, ld, nq, batch))) return rc;
    {
        dim3 grid(c->n / 256, nq + K), block(256);
#define FHE_KSP_CASE(BE) case BE: launch(k_ks_inner_ptr<BE, 2>, grid, block, s, c->T, nq, c->alpha, batch, acc, (const u64*)ext, \
                                         (const u64*)rlk, (const ConstF*)c->p_mod_q, mp); break;
        switch (beta) { FHE_KSP_CASE(1) FHE_KSP_CASE(2) FHE_KSP_CASE(3) default: launch(k_ks_inner_ptr<4, 2>, grid, block, s, c->T, nq,
                        c->alpha, batch, acc, (const u64*)ext, (const u64*)rlk, (const ConstF*)c->p_mod_q, mp); break; }
#undef FHE_KSP_CASE
    }
    return moddown_rescale_tail(c, s, (u64*)out, acc, nq, 2 * batch, "fhe_mul_relin_rescale_ptrs");
}

int fhe_ks_accum(fhe_ctx* c, void* stream, uint64_t* acc, const uint64_t* d, const uint64_t* ksk, const uint64_t* lift,
                 int lift_polys, int nq, int batch, int accumulate) {
    if (bad_shape(c, nq, 0) || c->n_p == 0 || batch < 1 || c->modup_beta[nq] > FHE_MAX_BETA || !acc ||
        (lift && lift_polys != 1 && lift_polys != 2) || (!d && !lift) || (d && !ksk))
        return fail("fhe_ks_accum: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    const int n = c->n, ne = nq + c->n_p, beta = c->modup_beta[nq];
    if (!d) {                                              // no key switch: acc (+)= P * lift
        launch_ks_inner(c, s, nq, batch, (u64*)acc, nullptr, (const u64*)lift, nullptr, (const u64*)lift, c->p_mod_q, false,
                        0, 0, lift_polys, accumulate, true);
        return check("fhe_ks_accum");
    }
    u64* base = arena(c, ((size_t)batch * nq + (size_t)batch * beta * ne) * n);
    if (!base) return fail("fhe_ks_accum: scratch allocation failed");
    u64* ext = base + (size_t)batch * nq * n;
    int rc;
    if ((rc = fhe_modup(c, stream, (uint64_t*)ext, d, nq, batch))) return rc;
    launch_ks_inner(c, s, nq, batch, (u64*)acc, (const u64*)ext, (const u64*)d, (const u64*)ksk, (const u64*)lift,
                    c->p_mod_q, false, 0, 0, lift_polys, accumulate);
    return check("fhe_ks_accum");
}

int fhe_moddown_rescale(fhe_ctx* c, void* stream, uint64_t* out, uint64_t* acc, int nq, int npoly) {
    if (bad_shape(c, nq, 0) || nq < 2 || c->n_p == 0 || npoly < 1 || !out || !acc)
        return fail("fhe_moddown_rescale: bad shape");
    return moddown_rescale_tail(c, (cudaStream_t)stream, (u64*)out, (u64*)acc, nq, npoly, "fhe_moddown_rescale");
}

int fhe_moddown(fhe_ctx* c, void* stream, uint64_t* out, uint64_t* acc, int nq, int npoly) {
    if (bad_shape(c, nq, 0) || c->n_p == 0 || npoly < 1) return fail("fhe_moddown: bad shape");
    cudaStream_t s = (cudaStream_t)stream;
    const int n = c->n, K = c->n_p, ne = nq + K;
    u64* accp = (u64*)acc + (size_t)nq * n;
    {   // special limbs -> coefficient domain, times (P/p_k)^-1, in place
        RowMap m = make_map(c, K, nq, nq);
        LoadPlain ld; ld.src = accp; ld.poly_stride = (long long)ne * n;
        StorePlain st; st.dst = accp; st.poly_stride = (long long)ne * n;
        ntt_inv(c, m, npoly * K, ld, accp, (long long)ne * n, st, c->moddown_scale, s);
    }
    // P -> q_i conversion written straight into `out` (coefficient domain) ...
    { const int ns = K; launch_bconv(c, s, &ns, npoly, c->moddown_table[nq], 1, (u64*)out, (long long)nq * n, accp, (long long)ne * n); }
    {   // ... then NTT it in place and finish: out = (acc - it) * P^-1
        RowMap m = make_map(c, nq, 0, nq);
        LoadPlain ld; ld.src = (const u64*)out; ld.poly_stride = (long long)nq * n;
        StoreSubMul st; st.out = (u64*)out; st.out_poly_stride = (long long)nq * n;
        st.in = (const u64*)acc; st.in_poly_stride = (long long)ne * n; st.c = c->pinv;
        ntt_fwd(c, m, npoly * nq, ld, (u64*)out, (long long)nq * n, st, s);
    }
    return check("fhe_moddown");
}

int fhe_keyswitch(fhe_ctx* c, void* stream, uint64_t* out, const uint64_t* d, const uint64_t* ksk, int nq, int batch) {
    if (bad_shape(c, nq, 0) || c->n_p == 0 || batch < 1) return fail("fhe_keyswitch: bad shape");
    const int n = c->n, ne = nq + c->n_p, beta = c->modup_beta[nq];
    // arena layout: [y : B*nq][ext : B*beta*ne][acc : 2*B*ne]   (fhe_modup uses the first B*nq rows)
    const size_t words = ((size_t)batch * nq + (size_t)batch * beta * ne + 2 * (size_t)batch * ne) * n;
    u64* base = arena(c, words);
    if (!base) return fail("fhe_keyswitch: scratch allocation failed");
    u64* ext = base + (size_t)batch * nq * n;
    u64* acc = ext + (size_t)batch * beta * ne * n;
    int rc;
    if ((rc = fhe_modup(c, stream, (uint64_t*)ext, d, nq, batch))) return rc;
    if ((rc = fhe_ks_inner(c, stream, (uint64_t*)acc, (const uint64_t*)ext, d, ksk, nq, batch))) return rc;
    return fhe_moddown(c, stream, out, (uint64_t*)acc, nq, 2 * batch);
}

int fhe_from_i64(fhe_ctx* c, void* stream, uint64_t* out, const int64_t* coeffs, int nq, int np, int batch) {
    if (bad_shape(c, nq, np) || batch < 1) return fail("fhe_from_i64: bad shape");
    const int rpp = nq + np;
    launch(k_from_i64, dim3(c->n / 256, batch * rpp), dim3(256), (cudaStream_t)stream, c->T, make_map(c, rpp, 0, nq),
           (u64*)out, (const long long*)coeffs);
    return check("fhe_from_i64");
}

int fhe_crt_centered(fhe_ctx* c, void* stream, double* out, const uint64_t* x, int limbs, int batch) {
    if (!c || limbs < 1 || limbs > 2 || limbs > c->n_q || batch < 1) return fail("fhe_crt_centered: bad shape");
    launch(k_crt_centered, dim3(c->n / 256, batch), dim3(256), (cudaStream_t)stream, c->T, out, (const u64*)x, limbs,
           c->q0inv_mod_q1);
    return check("fhe_crt_centered");
}

}  // extern "C"
