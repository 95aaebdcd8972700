/*
 * ckks_b200.h -- C ABI of libckks_b200.so, the sm_100a replacement for the reference's `ntt_cuda`
 * extension (Desilo/liberate-fhe, src/liberate/ntt/ntt.cpp + ntt_cuda_kernel.cu).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to int64 data on the current CUDA device unless noted;
 *   - a "[C,N] limb matrix" is C rows of N contiguous int64 coefficients, rows `stride` ELEMENTS apart
 *     (so the reference's strided row views d[:-K], x[start:] are passed as pointer + stride);
 *   - per-limb constant arrays (_2q, ql, qh, kl, kh, Rs, Ninv, ...) are contiguous int64[C] -- exactly the
 *     tensors ntt_context hands to ntt_cuda (src/liberate/ntt/ntt_context.py:138-189);
 *   - `stream` is a cudaStream_t (pass torch.cuda.current_stream().cuda_stream); launches are asynchronous;
 *   - return value: 0 on success, otherwise the cudaError_t of the failed launch, or a negative CKKS_E_*
 *     code for an argument the kernels do not support.  The reference checks nothing (ntt.cpp has no
 *     TORCH_CHECK) -- success-path behaviour is identical, failures are reported instead of ignored.
 *   - results are BIT-IDENTICAL to the reference kernels, including lazy [0,2q) representatives, for
 *     |inputs| < 2^62 (the reference's own arithmetic overflows beyond that).
 *
 * Each entry point cites the reference interface it replaces (paths relative to src/liberate/ntt/).
 */
#ifndef CKKS_B200_H
#define CKKS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CKKS_E_BADARG (-1)     /* null pointer / non-positive size */
#define CKKS_E_LOGN (-2)       /* logN outside [12, 17] */
#define CKKS_E_ALIGN (-3)      /* pointer or stride not 16-byte aligned */

int ckks_abi_version(void);                 /* 4; the Python loader refuses any other value */
/* tuning knobs (defaults are the measured best):
 *   2 = L2 prefetch distance in rows (28, 0 = off);  9 = MB of extended rows per key-switch slab (100);
 *   10 = internal side streams (2; 1..4);  11 = MB per row slab of a big batched transform (24);
 *   12 = rescale fused into the tensor stage's column pass (1);  16 = the two ModDown tails on two streams (1);
 *   17 = block passes read the last-group twiddles from the packed tables (1);
 *   18 = the executor keeps NTT-domain data in warp-interleaved order (1; needs permuted key copies, ckks_perm_rows);
 *   19 = hot-path kernels launched with programmatic stream serialization (1): the next grid ramps up under the tail of
 *        the previous one; every such kernel waits (griddepcontrol.wait) before its first global access.
 *   22 = least number of slabs of the key switch's forward chain (2): a 1/8 limb shard would fit one slab and run its
 *        kernels strictly one after the other; two slabs on two internal streams overlap the FP64 rows with the 60-bit rows.
 * Unknown keys return CKKS_E_BADARG.  The library is single-threaded per device (one host thread per device issues calls). */
int ckks_set_option(int key, int value);
int ckks_get_option(int key);               /* current value of a knob (negative: unknown key) */
int64_t ckks_launch_count(void);            /* kernels launched by the library since it was loaded */

/* ---- level 1: the 15 ntt_cuda operators (ntt.cpp:421-437), one device per call ---------------------- */

/* mont_mult (ntt.cpp:120-139, kern.cu:66-146): c = a (*) b, lazy Montgomery product, out of place */
int ckks_mont_mult(const int64_t* a, int64_t a_stride, const int64_t* b, int64_t b_stride, int64_t* c,
                   int64_t c_stride, int C, int N, const int64_t* ql, const int64_t* qh, const int64_t* kl,
                   const int64_t* kh, void* stream);

/* mont_enter (ntt.cpp:141-156, kern.cu:154-226): a[i][:] = mont(a[i][:], Rs[i]) in place (any per-limb scalar) */
int ckks_mont_enter(int64_t* a, int64_t a_stride, const int64_t* Rs, int C, int N, const int64_t* ql,
                    const int64_t* qh, const int64_t* kl, const int64_t* kh, void* stream);

/* ntt / enter_ntt (ntt.cpp:158-204, kern.cu:278-423): forward negacyclic NTT in place, natural -> bit-reversed.
 * tw = COMPACT Montgomery twiddles psi^bitrev(i), [C][N], rows tw_stride apart (build once with
 * ckks_compact_twiddles from the reference's painted psi[C][logN][N/2]).  Rs != NULL fuses mont_enter. */
int ckks_ntt(int64_t* a, int64_t a_stride, int C, int logN, const int64_t* tw, int64_t tw_stride,
             const int64_t* Rs, const int64_t* _2q, const int64_t* ql, const int64_t* qh, const int64_t* kl,
             const int64_t* kh, void* stream);

/* intt / intt_exit / intt_exit_reduce / intt_exit_reduce_signed (ntt.cpp:206-318, kern.cu:476-973):
 * inverse NTT in place, bit-reversed -> natural, x N^-1 (Ninv = N^-1 R mod q), then exit_mode
 * 0: nothing, 1: mont_redc, 2: + reduce to [0,q), 3: + make_signed. */
int ckks_intt(int64_t* a, int64_t a_stride, int C, int logN, const int64_t* tw, int64_t tw_stride,
              const int64_t* Ninv, const int64_t* _2q, const int64_t* ql, const int64_t* qh, const int64_t* kl,
              const int64_t* kh, int exit_mode, void* stream);

/* mont_redc (ntt.cpp:320-334, kern.cu:559-653) */
int ckks_mont_redc(int64_t* a, int64_t a_stride, int C, int N, const int64_t* ql, const int64_t* qh,
                   const int64_t* kl, const int64_t* kh, void* stream);
/* reduce_2q / make_signed / make_unsigned (ntt.cpp:336-376, kern.cu:664-699, 980-995) */
int ckks_reduce_2q(int64_t* a, int64_t a_stride, int C, int N, const int64_t* _2q, void* stream);
int ckks_make_signed(int64_t* a, int64_t a_stride, int C, int N, const int64_t* _2q, void* stream);
int ckks_make_unsigned(int64_t* a, int64_t a_stride, int C, int N, const int64_t* _2q, void* stream);
/* mont_add / mont_sub (ntt.cpp:378-407, kern.cu:1016-1058): out of place, lazy mod 2q */
int ckks_mont_add(const int64_t* a, int64_t a_stride, const int64_t* b, int64_t b_stride, int64_t* c,
                  int64_t c_stride, int C, int N, const int64_t* _2q, void* stream);
int ckks_mont_sub(const int64_t* a, int64_t a_stride, const int64_t* b, int64_t b_stride, int64_t* c,
                  int64_t c_stride, int C, int N, const int64_t* _2q, void* stream);
/* cc_add / cc_sub (engine.py:1268-1330): mont_add | mont_sub followed by reduce_2q, one pass instead of two (c may alias a or b) */
int ckks_addsub_reduce(const int64_t* a, int64_t a_stride, const int64_t* b, int64_t b_stride, int64_t* c,
                       int64_t c_stride, int C, int N, const int64_t* _2q, int sub, void* stream);
/* tile_unsigned (ntt.cpp:409-419, kern.cu:997-1014): dst[i][:] = a[:] + q_i */
int ckks_tile_unsigned(const int64_t* a, int64_t* dst, int64_t dst_stride, int C, int N, const int64_t* _2q,
                       void* stream);

/* painted psi[C][logN][N/2] (ckks_context.py:336-341) -> compact [C][N]; forward != 0 for psi, 0 for psi^-1 */
int ckks_compact_twiddles(const int64_t* painted, int64_t* compact, int C, int logN, int forward, void* stream);

/* ---- canonical-output ("fast") transforms for the fused path (additions) ----------------
 * Congruent to ntt/enter_ntt resp. intt* but with CANONICAL output: ckks_ntt_fast == {[x scal], ntt, reduce to [0,q)},
 * ckks_intt_fast == {intt stages, x scal, reduce to [0,q) or centred}.  Plain (non-Montgomery) twiddle tables:
 * tw_u64 [period][N] of {w, floor(w 2^64/q)} and tw_f64 [period][N] doubles (built by ckks_fast_tables from the
 * canonical plain table).  Row r uses the constants of limb r % period (batched key-switch extension).
 * Inputs: forward [0, 2q) (any |x| < 2^51 for primes < 2^42); inverse [0, 2q).  scal/scal_sh: per-limb plain
 * multiplier s and floor(s 2^64/q) (forward: optional, applied on load; inverse: required, e.g. N^-1 R^-1). */
int ckks_fast_tables(const int64_t* plain, const int64_t* q, void* tw_u64, double* tw_f64, int C, int N, void* stream);
/* PACKED copies of the last four stages of the fast tables (either table may be NULL): one thread of a block pass owns
 * 15 twiddles there; per 512-coefficient warp tile they are stored lane-interleaved so that the loads of a warp are
 * contiguous 256/512-byte runs (ntt_fast.cuh: TwPacked).  Same sizes as the plain tables: [C][N] x 16 B and [C][N] x 8 B. */
int ckks_fast_pack(const void* tw_u64, const double* tw_f64, void* twp_u64, double* twp_f64, int C, int logN, void* stream);
/* twp_u64 / twp_f64: the packed tables or NULL; qinv: [period] doubles 1/q or NULL (then computed per CTA);
 * perm != 0: the NTT-domain side (forward output / inverse input) is in warp-interleaved order (see ckks_perm_rows) */
int ckks_ntt_fast(int64_t* a, int64_t a_stride, int rows, int period, int logN, const void* tw_u64,
                  const double* tw_f64, const void* twp_u64, const double* twp_f64, const int64_t* q, const double* qinv,
                  const int64_t* scal, const uint64_t* scal_sh, int force_int, int perm, void* stream);
int ckks_intt_fast(int64_t* a, int64_t a_stride, int rows, int period, int logN, const void* tw_u64,
                   const double* tw_f64, const void* twp_u64, const double* twp_f64, const int64_t* q, const double* qinv,
                   const int64_t* scal, const uint64_t* scal_sh, int centred, int force_int, int perm, void* stream);
/* NTT-domain rows between natural order and the executor's warp-interleaved order (inverse != 0: back to natural):
 * inside every 512-coefficient tile, coefficient 16 t + k <-> position ((k >> 1) * 32 + t) * 2 + (k & 1).  Out of place.
 * Used once per evaluation / rotation key (the engine caches the permuted copy). */
int ckks_perm_rows(const int64_t* in, int64_t in_stride, int64_t* out, int64_t out_stride, int rows, int N, int inverse,
                   void* stream);

/* ---- level 2: fused hot-path operators (additions; engine.py line numbers = src/liberate/fhe/ckks_engine.py) */

/* rescale (engine.py:1026-1038): out[i] = reduce_q( mont(in[i] - r0, scale[i]) + (r0 > round_at) ).
 * `in` rows are the limbs that survive; r0 is the dropped limb (one row of N).  canon != 0 additionally maps the
 * (congruent) slightly negative representatives the reference leaves here into [0,q) -- fused path only. */
int ckks_rescale(const int64_t* in, int64_t in_stride, const int64_t* r0, int64_t* out, int64_t out_stride, int C,
                 int N, const int64_t* scale, int64_t round_at, int canon, const int64_t* _2q, const int64_t* ql,
                 const int64_t* qh, const int64_t* kl, const int64_t* kh, void* stream);
/* rescale with a per-limb Montgomery scalar folded in before and / or after it, one pass, the integers of the reference's
 * kernel sequences:  pre  != NULL: in[i] <- reduce_2q(mont(in[i], pre[i])) first  (mult_scalar, engine.py:2052-2098; r0 is
 * the dropped limb after the same scaling);  post != NULL: out[i] <- reduce_2q(mont(out[i], post[i])) last  (level_up,
 * engine.py:1410-1467).  Both NULL == ckks_rescale with canon = 0. */
int ckks_rescale_scaled(const int64_t* in, int64_t in_stride, const int64_t* r0, int64_t* out, int64_t out_stride, int C,
                        int N, const int64_t* pre, const int64_t* scale, int64_t round_at, const int64_t* post,
                        const int64_t* _2q, const int64_t* ql, const int64_t* qh, const int64_t* kl, const int64_t* kh,
                        void* stream);
/* plaintext x ciphertext in the NTT domain (mc_mult / pc_mult, engine.py:2100-2140: two mont_mult calls):
 * d0 = mont(p, c0), d1 = mont(p, c1); p, c0, c1 share in_stride; d0 / d1 may alias c0 / c1. */
int ckks_pc_product(const int64_t* p, const int64_t* c0, const int64_t* c1, int64_t in_stride, int64_t* d0, int64_t* d1,
                    int64_t out_stride, int C, int N, const int64_t* _2q, const int64_t* ql, const int64_t* qh,
                    const int64_t* kl, const int64_t* kh, void* stream);
/* plaintext + ciphertext (mc_add, engine.py:2142-2175): out = reduce_2q(mont_redc(mont_add(mont(p, Rs_scale), mont(c0, Rs)))),
 * the integers of the reference's five kernels, one pass.  Rs_scale = R^2 * 2^scale_bits mod q (nctx.py:138-142). */
int ckks_pc_add(const int64_t* p, int64_t p_stride, const int64_t* c0, int64_t c_stride, int64_t* out, int64_t out_stride,
                int C, int N, const int64_t* Rs_scale, const int64_t* Rs, const int64_t* _2q, const int64_t* ql,
                const int64_t* qh, const int64_t* kl, const int64_t* kh, void* stream);

/* tensor product (engine.py:1095-1101): d0 = x0*y0, d1 = x0*y1 (+) x1*y0, d2 = x1*y1 (lazy, NTT domain) */
int ckks_tensor_product(const int64_t* x0, const int64_t* x1, const int64_t* y0, const int64_t* y1, int64_t in_stride,
                        int64_t* d0, int64_t* d1, int64_t* d2, int64_t out_stride, int C, int N, const int64_t* _2q,
                        const int64_t* ql, const int64_t* qh, const int64_t* kl, const int64_t* kh, void* stream);

/* Garner mixed-radix digits of one partition (pre_extend, engine.py:654-705).
 * a: [alpha,N] rows of the partition (plain, as the reference reads them); state: [alpha,N] out (may alias a).
 * Y_scalar[alpha-1], L_scalar packed row-major upper triangle: entry (i, j) for step i (0..alpha-3) and target
 * row j (i+2..alpha-1) at  Ltri[i*alpha + j]  (ntt_context.py:336-349).  Per-limb constants of the alpha rows. */
int ckks_garner_digits(const int64_t* a, int64_t a_stride, int64_t* state, int64_t state_stride, int alpha, int N,
                       const int64_t* Y_scalar, const int64_t* Ltri, const int64_t* ql, const int64_t* qh,
                       const int64_t* kl, const int64_t* kh, void* stream);

/* basis extension of one partition's digits to E target limbs (extend, engine.py:707-743):
 * out[t] = mont(state[0], Rs[t]) (+) mont(state[1], Lenter[0][t]) (+) ...  (lazy mont_add chain, Montgomery form).
 * Lenter: [(alpha-1)][E] row-major.  Target-limb constants have length E.  canon != 0: result shifted into
 * [0, 2q) when the lazy chain ends below zero (fused path only). */
int ckks_extend(const int64_t* state, int64_t state_stride, int alpha, int64_t* out, int64_t out_stride, int E, int N,
                const int64_t* Rs, const int64_t* Lenter, int canon, const int64_t* _2q, const int64_t* ql,
                const int64_t* qh, const int64_t* kl, const int64_t* kh, void* stream);

/* evaluation-key inner product (switcher_later_part + the Sigma over parts, engine.py:906-937, 832-840):
 * acc0[t] (+)= mont(ext[t], ksk0[t]), acc1[t] (+)= mont(ext[t], ksk1[t]); first != 0 overwrites instead. */
int ckks_ksk_accumulate(const int64_t* ext, int64_t ext_stride, const int64_t* ksk0, const int64_t* ksk1,
                        int64_t ksk_stride, int64_t* acc0, int64_t* acc1, int64_t acc_stride, int E, int N, int first,
                        const int64_t* _2q, const int64_t* ql, const int64_t* qh, const int64_t* kl,
                        const int64_t* kh, void* stream);

/* the same inner product over ALL `parts` partitions in one pass (ext: [parts*E rows]; k0_ptrs/k1_ptrs: device
 * arrays of `parts` pointers to row 0 of each partition's key half, rows ksk_stride apart).  Congruent to the
 * running sum above (same Montgomery products, same summation order). */
int ckks_ksk_inner(const int64_t* ext, int64_t ext_stride, int parts, const int64_t* const* k0_ptrs,
                   const int64_t* const* k1_ptrs, int64_t ksk_stride, int64_t* acc0, int64_t* acc1, int64_t acc_stride,
                   int E, int N, const int64_t* _2q, const int64_t* ql, const int64_t* qh, const int64_t* kl,
                   const int64_t* kh, void* stream);

/* ModDown (engine.py:851-901): d is [E,N] = ordinary limbs (L rows) then K special rows, all plain in [0,q)
 * (after intt_exit_reduce).  Runs the K sequential "subtract special limb, multiply by P_j^-1" steps with the
 * reference's exact lazy semantics, then mont_redc + reduce on the ordinary rows.  PiR: [K][E] row-major
 * (engine.py:183-216; entries beyond the live rows of a step are ignored).  d is not modified; the
 * result goes to out; if add != NULL the result is (add + result) reduced to [0,q) (relinearize / switch_key tail,
 * engine.py:1135-1140, 947-948) written to out.  eff: caller-provided workspace of K*N int64. */
int ckks_moddown(int64_t* d, int64_t d_stride, int L, int K, int N, const int64_t* Rs, const int64_t* PiR,
                 const int64_t* add, int64_t add_stride, int64_t* out, int64_t out_stride, int64_t* eff,
                 const int64_t* _2q, const int64_t* ql, const int64_t* qh, const int64_t* kl, const int64_t* kh,
                 void* stream);

/* Galois automorphism of coefficient rows (encdec.rotate / conjugate, encdec.py:224-270) fused with
 * make_unsigned + reduce_2q (engine.py:1196-1200) when canon != 0:
 * out[i][(g*j) mod N] = +-in[i][j]  (sign = -1 when (g*j mod 2N) >= N), g odd. */
int ckks_automorphism(const int64_t* in, int64_t in_stride, int64_t* out, int64_t out_stride, int C, int N,
                      int64_t g, int canon, const int64_t* _2q, void* stream);

/* ---- level 3: fused executor -- one C call runs a whole stage of the hot path (all kernels on `stream`) -------
 * ckks_level_t describes one (level, device): every pointer is a DEVICE pointer the caller keeps alive.
 * Rows = the device's limbs live at this level: L ordinary rows then K special rows (E = L + K). */
typedef struct {
    int32_t logN, L, K, nparts, nlocal, _pad;
    const int64_t *q, *_2q, *ql, *qh, *kl, *kh, *Rs;          /* [E] per-row constants                          */
    const void* twf_u64; const double* twf_f64;               /* fast forward tables of the E rows              */
    const void* twi_u64; const double* twi_f64;               /* fast inverse tables                            */
    const int64_t *sR, *sR_sh, *sExit, *sExit_sh;             /* [E] plain scalars R and N^-1 R^-1 (+ Shoup)    */
    const int64_t* PiR;                                       /* [K][E] ModDown table                           */
    const int32_t* part_alpha;                                /* [nparts] rows of every partition (storage order) */
    const int64_t* const* Lenter;                             /* [nparts] -> [(alpha-1)][E]                     */
    const int32_t *loc_row0, *loc_alpha;                      /* [nlocal] partitions whose digits this device makes */
    const int64_t* const* loc_Y; const int64_t* const* loc_Ltri;
    const int64_t* rescale_scale; int64_t round_at;           /* rescale INTO this level: [L] multipliers       */
    /* optional (NULL = separate extend kernel): tables of the extension fused into the forward column pass */
    const int32_t* part_wide;                                 /* [nparts] 1: digits may exceed 2^51 (alpha == 1) */
    const double* const* Hm;                                  /* [nparts] -> [(alpha-1)][E] doubles m_i mod q_t  */
    const double *Rd, *C31;                                   /* [E] R mod q_t, 2^31 mod q_t as doubles          */
    const double* Rinv;                                       /* [E] R^-1 mod q_t (FP64 inner product), or NULL  */
    const double* Pinv;                                       /* [K][E] P_i^-1 mod q_t (FP64 ModDown), or NULL   */
    int32_t L_small, amax;                                    /* leading ordinary rows with q < 2^42; max alpha  */
    const void* twpf_u64; const double* twpf_f64;             /* packed forward tables (ckks_fast_pack) or NULL  */
    const void* twpi_u64; const double* twpi_f64;             /* packed inverse tables or NULL                   */
    const double* qinv;                                       /* [E] 1/q_t as doubles, or NULL                   */
} ckks_level_t;

/* rescale x4 -> batched enter+NTT -> tensor product -> batched iNTT+exit -> Garner digits of d2
 * (cc_mult + the first half of relinearize, engine.py:1072-1129, 654-705).  a0..b1: the rows that survive the rescale
 * ([L][N], in_stride); r0*: the dropped limb of each polynomial ([N], on this device).  x: workspace [4][L][N];
 * d: out [3][L][N] plain canonical d0,d1,d2; digits: out [L][N] (rows of the local partitions). */
int ckks_exec_tensor_stage(const ckks_level_t* lv, const int64_t* a0, const int64_t* a1, const int64_t* b0,
                           const int64_t* b1, int64_t in_stride, const int64_t* r0a0, const int64_t* r0a1,
                           const int64_t* r0b0, const int64_t* r0b1, int64_t* x, int64_t* d, int64_t* digits,
                           void* stream);
/* Garner digits of the local partitions of a [L][N] polynomial (pre_extend for every partition, one launch).
 * galois = 0, or an odd g < 2N: the digits of the Galois image of a (encdec.py:224-246 + engine.py:1196-1200: out[(g j) mod N]
 * = +-a[j], canonical) are computed by gathering from the UNROTATED canonical rows (a != digits then).
 * CKKS_E_BADARG when a partition has more than 8 limbs (lv->amax). */
int ckks_exec_digits(const ckks_level_t* lv, const int64_t* a, int64_t a_stride, int64_t* digits, int64_t d_stride,
                     int64_t galois, void* stream);
/* extend (all partitions) -> batched NTT -> evk inner product -> batched iNTT+exit -> ModDown (+add, reduce)
 * (create_switcher engine.py:812-904 + the relinearize / switch_key tails).  digit_ptrs: device [nparts] pointers to
 * each partition's [alpha][N] digit block (rows digit_stride apart) -- local or received from a peer;
 * k0_ptrs / k1_ptrs: device [nparts] row-0 pointers of the key halves; keys_permuted != 0: they point to copies made
 * with ckks_perm_rows (the stage then keeps its NTT-domain data in the same order);
 * add0_galois = 0, or an odd g < 2N: the polynomial added to output 0 is the Galois image of add0 (rotate_single: the
 * rotated c0, engine.py:1194-1200, 947), gathered from the unrotated canonical rows inside the ModDown kernel;
 * ws: ckks_exec_keyswitch_ws_elems(...) int64 elements;
 * phase: 3 = the whole stage; 1 = only extend + batched NTT (the NTT-domain extended block stays in ws); 2 = only inner
 * product + inverse NTT + ModDown on the block a phase-1 call left in ws -- hoisted rotations share one phase-1 call among
 * many keys (engine.rotate_hoisted; keys_permuted must be the same in both calls);
 * part_begin, part_end: phase 1 only -- the range of partitions (positions in lv's partition order) to extend and transform,
 * -1, -1 = all: one process per GPU transforms its own partitions while the peers' digit blocks are still arriving. */
int ckks_exec_keyswitch_stage(const ckks_level_t* lv, const int64_t* const* digit_ptrs, int64_t digit_stride,
                              const int64_t* const* k0_ptrs, const int64_t* const* k1_ptrs, int64_t ksk_stride,
                              int keys_permuted, const int64_t* add0, const int64_t* add1, int64_t add_stride,
                              int64_t add0_galois, int64_t* out0, int64_t* out1, int64_t out_stride, int64_t* ws,
                              int phase, int part_begin, int part_end, void* stream);
int64_t ckks_exec_keyswitch_ws_elems(int L, int K, int nparts, int N);

/* ---- sampler: ChaCha20 counter mode (replaces src/liberate/csprng/: chacha20 / randint / discrete_gaussian / randround
 * extensions, csprng.py:18-323).  Same stream as the reference for the same key and nonce, but stateless: the block
 * of channel c, position l has the 64-bit counter ctr_base[c] + l + epoch[c][l] * inc, and epoch[c][l] (device, uint32,
 * incremented by the call) replaces the reference's [blocks,16] int64 state tensor.  One block gives four samples, so a
 * channel of N coefficients has L = N/4 blocks.  key_nonce: HOST pointer to 8 key words + 2 nonce words.
 * out pointers must be 32-byte aligned. */
/* randbytes (csprng.py:216-236, chacha20_cuda_kernel.cu): out[C][L][16] = the 16 output words of every block */
int ckks_rng_bytes(int64_t* out, int C, int L, const uint32_t* key_nonce, const uint64_t* ctr_base, uint32_t* epoch,
                   uint64_t inc, void* stream);
/* randint (csprng.py:238-272, randint_cuda_kernel.cu:23-102): out[C][4L] = floor(X q[c] / 2^128) + shift, X the 128-bit draw */
int ckks_rng_randint(int64_t* out, int C, int L, const uint64_t* q, int64_t shift, const uint32_t* key_nonce,
                     const uint64_t* ctr_base, uint32_t* epoch, uint64_t inc, void* stream);
/* discrete_gaussian (csprng.py:274-305, discrete_gaussian_cuda_kernel.cu:27-108): out[C][4L]; lut: HOST table of the CDT
 * search tree, lut_size low words then lut_size high words (discrete_gaussian_sampler.py:96-116), depth levels */
int ckks_rng_gaussian(int64_t* out, int C, int L, const uint64_t* lut, int lut_size, int depth, const uint32_t* key_nonce,
                      const uint64_t* ctr_base, uint32_t* epoch, uint64_t inc, void* stream);
/* randround (csprng.py:307-323, randround_cuda_kernel.cu:8-37): out[i] = sign(coef[i]) (floor|coef[i]| + [u_i < frac 2^32]),
 * u_i the i-th 32-bit word of blocks 0..ceil(n/16)-1 of one channel (ctr_base has one entry) */
int ckks_rng_randround(const double* coef, int64_t* out, int n, const uint32_t* key_nonce, const uint64_t* ctr_base,
                       uint32_t* epoch, uint64_t inc, void* stream);

#ifdef __cplusplus
}
#endif
#endif
