/* aesfhe_b200.h -- C ABI of the B200 CKKS evaluation backend (libaesfhe_b200.so).
 *
 * This is the drop-in boundary below the reference's Python services.  The reference
 * (songhayeong/aes-fhe) reaches its ciphertext arithmetic only through the Python API of
 * the third-party module `desilofhe` (Engine / Ciphertext / Plaintext; import sites
 * engine_context.py:6, xor_service.py:12,69, gf_service.py:7, new.py:6).  A maintainer
 * binds these entry points with ctypes; every function below names the
 * `desilofhe.Engine` member(s) whose arithmetic it replaces and the reference call sites.
 *
 * Conventions
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers to 64-bit residues
 *     unless stated otherwise; `stream` is a cudaStream_t passed as void*.
 *   - a polynomial block is [nq + np][N] u64: the first nq active q-limbs (moduli 0..nq-1)
 *     followed by np special limbs (np is 0 or the context's n_p); blocks of one call are
 *     contiguous: [npoly][nq + np][N].  Residues are canonical [0,q), NTT domain
 *     (bit-reversed spectrum) unless a function says "coefficient domain".
 *   - return value 0 = ok; negative = error, text via fhe_last_error().  No function
 *     synchronises the stream (fhe_ntt_fused_status and fhe_ctx_destroy excepted); nothing is allocated per
 *     call except when the context's scratch arena has to grow (first call at a new size).
 *   - ONE STREAM PER CONTEXT: the scratch arena and the hand-over counters of the single-launch NTTs belong to
 *     the context, so all calls on one context must be issued on one stream (or be serialised by the caller).
 *     Use one context per stream for concurrency.  The Python host layer enforces this (backend_cuda.py).
 *   - the single-launch NTTs never hang: a hand-over wait that times out sets a sticky error flag and the
 *     transform's output is then wrong.  Callers MUST check fhe_ntt_fused_status() before trusting results that
 *     leave the device (the host layer does so on every decrypt / download / synchronize).
 */
#ifndef AESFHE_B200_H
#define AESFHE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fhe_ctx fhe_ctx;

/* Engine(...) construction (engine_context.py:32-56): builds twiddle / base-conversion
 * tables on `device` for the RNS chain `moduli` = q_0..q_{n_q-1}, p_0..p_{n_p-1} with
 * primitive 2N-th roots `psi` (host arrays of n_q + n_p entries); alpha = q-limbs per
 * key-switch digit. */
int fhe_ctx_create(fhe_ctx** out, int log_n, int n_q, int n_p, int alpha,
                   const uint64_t* moduli, const uint64_t* psi, int device);
void fhe_ctx_destroy(fhe_ctx* ctx);
const char* fhe_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
uint64_t fhe_launch_count(void);
/* number of length-N transforms (forward + inverse NTT rows) launched since load: the work unit bench.py uses to
 * scale the CPU sample of the reference arm to a whole AES-128 (the schedule's row count does not depend on N) */
uint64_t fhe_ntt_row_count(void);

/* NTT variant used by every entry point below (results are bit-identical):
 *   2 (default)  csrc/ntt_chained.cuh -- both passes in ONE launch, ticket-ordered so that the lazy
 *                intermediate is consumed from L2 (1.2 MB of DRAM traffic per limb instead of 2.0);
 *   0            csrc/ntt.cuh -- two launches per transform;
 *   1            csrc/ntt_fused.cuh -- persistent cooperative kernel (first enable calibrates the groups).
 * fhe_set_ntt_fused(ctx, mode), or env FHE_NTT_FUSED=mode at context creation.
 * fhe_ntt_fused_status synchronises the device and returns non-zero if a hand-over wait of modes 1 / 2
 * ever timed out (they never hang). */
int fhe_set_ntt_fused(fhe_ctx* ctx, int enabled);
int fhe_ntt_fused_status(fhe_ctx* ctx);

/* Forward / inverse negacyclic NTT, in place.  Used by encode/encrypt/decrypt
 * (xor_service.py:59-66) and inside every multiply / rotate / conjugate. */
int fhe_ntt_fwd(fhe_ctx* ctx, void* stream, uint64_t* data, int npoly, int nq, int np);
int fhe_ntt_inv(fhe_ctx* ctx, void* stream, uint64_t* data, int npoly, int nq, int np);

/* Engine.add(ct, ct) / subtract / pointwise products (xor_service.py:76,
 * sbox/sbox_service.py:105,112,129,136).  out, a: [npoly][batch][nq+np][N];
 * b: [b_npoly][b_batch][nq+np][N] with b_npoly in {npoly, 1} and b_batch in {batch, 1}
 * (1 = broadcast: plaintext masks and keys are shared by the polys / the batch). */
int fhe_add(fhe_ctx* ctx, void* stream, uint64_t* out, const uint64_t* a, const uint64_t* b,
            int npoly, int batch, int b_npoly, int b_batch, int nq, int np);
int fhe_sub(fhe_ctx* ctx, void* stream, uint64_t* out, const uint64_t* a, const uint64_t* b,
            int npoly, int batch, int b_npoly, int b_batch, int nq, int np);
int fhe_mul(fhe_ctx* ctx, void* stream, uint64_t* out, const uint64_t* a, const uint64_t* b,
            int npoly, int batch, int b_npoly, int b_batch, int nq, int np);
int fhe_neg(fhe_ctx* ctx, void* stream, uint64_t* out, const uint64_t* a, int npoly, int nq, int np);

/* Engine.multiply(ct, ct) tensor step (xor_service.py:71, sbox/sbox_service.py:114):
 * out[3][batch][nq][N] = (a0 b0, a0 b1 + a1 b0, a1 b1) for a, b = [2][batch][nq][N]. */
int fhe_tensor(fhe_ctx* ctx, void* stream, uint64_t* out, const uint64_t* a, const uint64_t* b, int nq, int batch);

/* Engine.multiply(ct, float|const Plaintext) / add_plain (xor_service.py:78-83,98,282-285;
 * sbox/sbox_service.py:124-136): multiply (or add) by the encoding of a complex constant,
 * given per limb slot as host arrays c_first[j], c_second[j] (value on the first / second
 * half of the bit-reversed spectrum). */
int fhe_mul_const(fhe_ctx* ctx, void* stream, uint64_t* out, const uint64_t* a,
                  const uint64_t* c_first, const uint64_t* c_second, int npoly, int nq, int np);
int fhe_add_const(fhe_ctx* ctx, void* stream, uint64_t* out, const uint64_t* a,
                  const uint64_t* c_first, const uint64_t* c_second, int npoly, int nq, int np);

/* LUT inner sums (the multiply(term, pt) + add loops of xor_service.py:283-285 and
 * sbox/sbox_service.py:124-136, batched): out_m = sum_t c[m][t] (.) in_t (+ c0[m] on poly 0).
 *   in[t]   : T (<= 16) device pointers to ciphertexts [2][batch][in_nq[t]][N], in_nq[t] >= nq
 *             (host array of pointers; higher-level inputs are read on their first nq limbs)
 *   consts  : device, [M][T][nq][2] pairs (c, RN(c/q_j)) of doubles -- the constant's residue on
 *             the first / second half of the bit-reversed spectrum
 *   c0      : device, [M][nq][2] residues added to polynomial 0, or NULL
 *   out     : M ciphertexts [2][batch][nq][N], contiguous.  No rescale is performed. */
int fhe_lincomb(fhe_ctx* ctx, void* stream, uint64_t* out, const uint64_t* const* in, const int* in_nq,
                const double* consts, const uint64_t* c0, int M, int T, int nq, int batch,
                const long long* in_poly_stride);
/* in_poly_stride (fhe_lincomb) / a_poly_stride (fhe_tensor_acc): NULL, or per input the distance in words between its
 * two polynomials -- an input may then be a batch slice of a larger tensor [2][B][nq_t][N] (pointer to element
 * [0][lo], stride B * nq_t * N) and is read in place. */

/* Rotate-mask-add and diagonal-matrix sums (shiftrows_service.py:41-50; the linear transforms of
 * Engine.bootstrap): out[2][batch][nq][N] (+)= sum_t a_t (.) p_t for T <= 16 terms in one pass.
 * a[t]: ciphertexts [2][batch][a_nq[t] >= nq][N] (host array of device pointers), p[t]: plaintexts
 * [nq][N] (NTT domain) shared by both polynomials and the batch.  No rescale is performed. */
int fhe_mul_plain_sum(fhe_ctx* ctx, void* stream, uint64_t* out, const uint64_t* const* a, const int* a_nq,
                      const uint64_t* const* p, int T, int nq, int batch, int accumulate);

/* Double-hoisted baby steps of a baby-step/giant-step linear transform (the CoeffToSlot / SlotToCoeff matrices of
 * Engine.bootstrap, xor_service.py:120-129): for nb <= 16 baby rotations of ONE ciphertext batch and G <= 4 giant
 * steps,  out[G][2][batch][nq+K][N],  out_g = sum_b pts[g * nb + b] (.) ( <sigma_b(ext), keys[b]> + P (sigma_b(c0), 0) )
 * in the extended basis: no rotated ciphertext is stored and no ModDown happens per baby step (one per giant step,
 * by the caller: fhe_moddown / fhe_ks_accum / fhe_moddown_rescale).  ext = fhe_modup output of c1 of `ct`
 * ([2][batch][ct_nq >= nq][N]); keys[b] = Galois key of baby step b with Galois element galois[b] (host arrays of
 * device pointers / values), keys[b] == NULL = no rotation; pts[.] = plaintexts [nq+K][N] in the extended basis,
 * NULL = absent diagonal.  accumulate != 0: the sums are added to what `out` holds (a transform with more than 16 baby
 * steps is evaluated in several passes over the same `ext`). */
int fhe_bsgs_inner(fhe_ctx* ctx, void* stream, uint64_t* out, const uint64_t* ext, const uint64_t* ct, int ct_nq,
                   const uint64_t* const* keys, const uint64_t* galois, const uint64_t* const* pts, int nb, int G,
                   int nq, int batch, int accumulate);

/* The same pattern for G output sums over ONE set of T ciphertexts (the diagonal sums of all giant steps of a
 * baby-step/giant-step linear transform in Engine.bootstrap, xor_service.py:120-129): out[G][2][batch][nq][N],
 * out_g = sum_t a_t (.) p[g * T + t]; a NULL plaintext pointer skips that term.  Every a_t is read once. */
int fhe_mul_plain_multi(fhe_ctx* ctx, void* stream, uint64_t* out, const uint64_t* const* a, const int* a_nq,
                        const uint64_t* const* p, int T, int G, int nq, int batch);

/* Lazy-relinearised products: acc[3][batch][nq][N] (+)= sum_g a_g (x) b_g for G <= 16 products;
 * a[g] are ciphertexts [2][a_batch[g]][a_nq[g]][N] (host array of device pointers), b is G
 * contiguous ciphertexts [2][b_batch][nq][N]; a_batch[g] and b_batch are `batch` or 1 (a
 * single ciphertext, e.g. an encrypted round key, broadcast over the batch).  One
 * fhe_keyswitch then relinearises the sum. */
int fhe_tensor_acc(fhe_ctx* ctx, void* stream, uint64_t* acc, const uint64_t* const* a, const int* a_nq,
                   const int* a_batch, const uint64_t* b, int b_batch, int G, int nq, int batch, int accumulate,
                   const long long* a_poly_stride, long long acc_poly_stride, const uint64_t* init);
/* acc_poly_stride: 0, or the distance in words between the polynomials of acc (acc may be a batch slice of a larger
 * [3][B][nq][N] tensor, so that the sums of several LUT outputs land in ONE tensor and are relinearised together);
 * init: NULL, or a 2-polynomial term [2][batch][nq][N] added to polynomials 0 and 1 (the row of a LUT whose outer
 * factor is the constant 1). */

/* Rescale after every multiply: [npoly][nq][N] -> [npoly][nq-1][N], division by q_{nq-1}
 * rounded to nearest. */
int fhe_rescale(fhe_ctx* ctx, void* stream, uint64_t* out, const uint64_t* in, int npoly, int nq);

/* Engine.bootstrap, first step (xor_service.py:120-129): ModRaise.  in[npoly][1][N] (level 0)
 * -> out[npoly][nq_out][N]: the centred representative of every coefficient mod q_0, reduced
 * into the first nq_out moduli (the plaintext becomes m + q_0 * I with a small integer I). */
int fhe_mod_raise(fhe_ctx* ctx, void* stream, uint64_t* out, const uint64_t* in, int npoly, int nq_out);

/* Engine.rotate / conjugate permutation part (xor_service.py:89,105): X -> X^galois on
 * `nrows` NTT-domain rows. */
int fhe_automorphism(fhe_ctx* ctx, void* stream, uint64_t* out, const uint64_t* in,
                     uint64_t galois, int nrows);

/* Hybrid key switching (Engine.relinearize, the relin inside multiply(ct,ct,rlk),
 * rotate, conjugate).  ksk: [dnum][2][n_q + n_p][N] over the full chain.
 * A batch of `batch` independent ciphertexts shares one pass over the key.
 *   fhe_keyswitch : d[batch][nq][N] -> out[2][batch][nq][N]   (= moddown(inner(modup(d))))
 * The three phases are exported for hoisted rotations and for phase-level parity tests:
 *   fhe_modup    : d[batch][nq][N] -> ext[batch][beta][nq+n_p][N]   (rows of a digit's own
 *                  limbs are left untouched; the inner product reads them from d)
 *   fhe_ks_inner : (ext, d, ksk) -> acc[2][batch][nq+n_p][N]
 *   fhe_moddown  : acc[npoly][nq+n_p][N] -> out[npoly][nq][N]   (acc's special limbs are
 *                  overwritten with scratch data) */
int fhe_keyswitch(fhe_ctx* ctx, void* stream, uint64_t* out, const uint64_t* d,
                  const uint64_t* ksk, int nq, int batch);
int fhe_modup(fhe_ctx* ctx, void* stream, uint64_t* ext, const uint64_t* d, int nq, int batch);
int fhe_ks_inner(fhe_ctx* ctx, void* stream, uint64_t* acc, const uint64_t* ext,
                 const uint64_t* d, const uint64_t* ksk, int nq, int batch);
int fhe_moddown(fhe_ctx* ctx, void* stream, uint64_t* out, uint64_t* acc, int nq, int npoly);

/* Engine.multiply(ct, ct, rlk) tail (xor_service.py:71): relinearise and rescale in one step.
 * d3[3][batch][nq][N] (tensor product) -> out[2][batch][nq-1][N] =
 * round(((d0, d1) * P + <ModUp(d2), rlk>) / (P * q_{nq-1})): the ModDown by P and the rescale by
 * q_{nq-1} share one base conversion and one set of NTTs. */
int fhe_relin_rescale(fhe_ctx* ctx, void* stream, uint64_t* out, const uint64_t* d3, const uint64_t* rlk,
                      int nq, int batch);

/* Engine.multiply(ct, ct, rlk) in one call (xor_service.py:71, and every product of make_power_basis,
 * xor_service.py:86 / sbox/sbox_service.py:93): a[2][batch][a_nq][N], b[2][batch][b_nq][N] with
 * a_nq, b_nq >= nq (only limbs 0..nq-1 are used: an operand of a higher level is used as it lies)
 * -> out[2][batch][nq-1][N], the same residues as fhe_tensor followed by fhe_relin_rescale.  The
 * 3-polynomial product is never stored: the key switch's first inverse NTT forms d2 = a1 b1 on load
 * and the key inner product forms P (a0 b0, a0 b1 + a1 b0) from the operands. */
int fhe_mul_relin_rescale(fhe_ctx* ctx, void* stream, uint64_t* out, const uint64_t* a, int a_nq,
                          const uint64_t* b, int b_nq, const uint64_t* rlk, int nq, int batch);

/* The same product with GATHERED operands (the batched products of the AES-128 service multiply slices, permutations
 * and concatenations of its state tensors -- the reference's drivers loop over single ciphertexts,
 * test_all_process.py:21-48): batch element i of the product is (a0[i], a1[i]) x (b0[i], b1[i]); a0, a1, b0, b1 are
 * HOST arrays of `batch` device pointers, each to one polynomial of >= nq contiguous limbs [>= nq][N].
 * 1 <= batch <= 128, at most four key-switching digits at this level.  out[2][batch][nq-1][N]. */
int fhe_mul_relin_rescale_ptrs(fhe_ctx* ctx, void* stream, uint64_t* out, const uint64_t* const* a0,
                               const uint64_t* const* a1, const uint64_t* const* b0, const uint64_t* const* b1,
                               const uint64_t* rlk, int nq, int batch);

/* Key switches whose ModDown is shared (the giant steps of a baby-step/giant-step linear transform inside
 * Engine.bootstrap, xor_service.py:120-129): acc[2][batch][nq + K][N] (+)= <ModUp(d), ksk> + P * lift, all in the
 * extended basis.  d[1][batch][nq][N] (NTT domain) or NULL for "no key switch, lift only"; lift[lift_polys][batch][nq][N]
 * or NULL, lift_polys = 1 lifts polynomial 0 only (a rotation: (sigma(c0), 0)); accumulate != 0 adds to acc. */
int fhe_ks_accum(fhe_ctx* ctx, void* stream, uint64_t* acc, const uint64_t* d, const uint64_t* ksk,
                 const uint64_t* lift, int lift_polys, int nq, int batch, int accumulate);

/* ... and the shared ModDown, merged with the rescale that follows: acc[npoly][nq + K][N] (overwritten) ->
 * out[npoly][nq-1][N] = round(acc / (P * q_{nq-1})). */
int fhe_moddown_rescale(fhe_ctx* ctx, void* stream, uint64_t* out, uint64_t* acc, int nq, int npoly);

/* Engine.encode / encrypt residue step (xor_service.py:59-66): signed 64-bit coefficients
 * coeffs[batch][N] (device) -> coefficient-domain residues out[batch][nq + np][N]. */
int fhe_from_i64(fhe_ctx* ctx, void* stream, uint64_t* out, const int64_t* coeffs, int nq, int np, int batch);

/* Engine.decrypt tail (xor_service.py:62-63): centred CRT of coefficient-domain limbs
 * 0..limbs-1 (limbs = 1 or 2) of x[batch][limbs][N] to doubles out[batch][N]. */
int fhe_crt_centered(fhe_ctx* ctx, void* stream, double* out, const uint64_t* x, int limbs, int batch);

#ifdef __cplusplus
}
#endif
#endif
