/*
 * ckks_oracle.c -- CPU restatement of the reference's RNS-CKKS limb arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path in
 * liberate-fhe_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product never does.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/src/liberate):
 *   kern.cu = ntt/ntt_cuda_kernel.cu, cctx.py = fhe/context/ckks_context.py
 *
 * Arithmetic model: all values are int64 with two's-complement wrap-around
 * (compile with -fwrapv) and arithmetic right shift, exactly what the CUDA
 * int64 kernels of the reference do.  R = 2^62, 31-bit halves (kern.cu:19-23).
 *
 * Pinned against: (1) python-int exact mathematics, (2) tables produced by the
 * reference's own ckks_context (tests/golden), (3) the reference's CUDA kernels
 * built from /root/reference into oracle/_ref (GPU tests).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define NBITS 62
#define HALF 31
static const int64_t FB_MASK = (((int64_t)1) << NBITS) - 1;
static const int64_t LB_MASK = (((int64_t)1) << HALF) - 1;

/* kern.cu:12-59  mont_mult_scalar_cuda_kernel: lazy Montgomery product, no final subtract. */
static inline int64_t mm(int64_t a, int64_t b, int64_t ql, int64_t qh, int64_t kl, int64_t kh)
{
    const int64_t al = a & LB_MASK, ah = a >> HALF;
    const int64_t bl = b & LB_MASK, bh = b >> HALF;
    const int64_t alpha = ah * bh;
    const int64_t beta = ah * bl + al * bh;
    const int64_t gamma = al * bl;
    /* s = x k mod R */
    const int64_t gammal = gamma & LB_MASK, gammah = gamma >> HALF;
    const int64_t betal = beta & LB_MASK, betah = beta >> HALF;
    int64_t upper = gammal * kh;
    upper = upper + (gammah + betal) * kl;
    upper = (int64_t)((uint64_t)upper << HALF);
    int64_t s = upper + gammal * kl;
    s = s & FB_MASK;
    /* t = x + s q ; u = t / R */
    const int64_t sl = s & LB_MASK, sh = s >> HALF;
    const int64_t sqb = sh * ql + sl * qh;
    const int64_t sqbl = sqb & LB_MASK, sqbh = sqb >> HALF;
    int64_t carry = (gamma + sl * ql) >> HALF;
    carry = (carry + betal + sqbl) >> HALF;
    return alpha + betah + sqbh + carry + sh * qh;
}

int64_t orc_mont_mult_scalar(int64_t a, int64_t b, int64_t ql, int64_t qh, int64_t kl, int64_t kh)
{
    return mm(a, b, ql, qh, kl, kh);
}

/* kern.cu:559-607  mont_redc_cuda_kernel: x * R^-1, no final subtract. */
static inline int64_t redc(int64_t x, int64_t ql, int64_t qh, int64_t kl, int64_t kh)
{
    const int64_t xl = x & LB_MASK, xh = x >> HALF;
    const int64_t xkb = xh * kl + xl * kh;
    int64_t s = (int64_t)((uint64_t)xkb << HALF) + xl * kl;
    s = s & FB_MASK;
    const int64_t sl = s & LB_MASK, sh = s >> HALF;
    const int64_t sqb = sh * ql + sl * qh;
    const int64_t sqbl = sqb & LB_MASK, sqbh = sqb >> HALF;
    int64_t carry = (x + sl * ql) >> HALF;
    carry = (carry + sqbl) >> HALF;
    return sqbh + carry + sh * qh;
}

/* All batched ops take a row stride (in elements) so strided row views work like
 * the reference's PackedTensorAccessor32 indexing. */

/* kern.cu:66-90 */
void orc_mont_mult(const int64_t *a, ptrdiff_t as, const int64_t *b, ptrdiff_t bs, int64_t *c, ptrdiff_t cs,
                   int C, int N, const int64_t *ql, const int64_t *qh, const int64_t *kl, const int64_t *kh)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < N; ++j)
            c[i * cs + j] = mm(a[i * as + j], b[i * bs + j], ql[i], qh[i], kl[i], kh[i]);
}

/* kern.cu:154-177 */
void orc_mont_enter(int64_t *a, ptrdiff_t as, const int64_t *Rs, int C, int N,
                    const int64_t *ql, const int64_t *qh, const int64_t *kl, const int64_t *kh)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < N; ++j)
            a[i * as + j] = mm(a[i * as + j], Rs[i], ql[i], qh[i], kl[i], kh[i]);
}

/* kern.cu:559-607 */
void orc_mont_redc(int64_t *a, ptrdiff_t as, int C, int N,
                   const int64_t *ql, const int64_t *qh, const int64_t *kl, const int64_t *kh)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < N; ++j)
            a[i * as + j] = redc(a[i * as + j], ql[i], qh[i], kl[i], kh[i]);
}

/* kern.cu:664-680 */
void orc_reduce_2q(int64_t *a, ptrdiff_t as, int C, int N, const int64_t *_2q)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < N; ++j) {
            const int64_t q = _2q[i] >> 1, v = a[i * as + j];
            a[i * as + j] = (v < q) ? v : v - q;
        }
}

/* kern.cu:682-699 */
void orc_make_signed(int64_t *a, ptrdiff_t as, int C, int N, const int64_t *_2q)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < N; ++j) {
            const int64_t q = _2q[i] >> 1, qh = q >> 1, v = a[i * as + j];
            a[i * as + j] = (v <= qh) ? v : v - q;
        }
}

/* kern.cu:980-995 */
void orc_make_unsigned(int64_t *a, ptrdiff_t as, int C, int N, const int64_t *_2q)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < N; ++j)
            a[i * as + j] += _2q[i] >> 1;
}

/* kern.cu:997-1014 */
void orc_tile_unsigned(const int64_t *a, int64_t *dst, ptrdiff_t ds, int C, int N, const int64_t *_2q)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < N; ++j)
            dst[i * ds + j] = a[j] + (_2q[i] >> 1);
}

/* kern.cu:1016-1036 */
void orc_mont_add(const int64_t *a, ptrdiff_t as, const int64_t *b, ptrdiff_t bs, int64_t *c, ptrdiff_t cs,
                  int C, int N, const int64_t *_2q)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < N; ++j) {
            const int64_t s = a[i * as + j] + b[i * bs + j];
            c[i * cs + j] = (s < _2q[i]) ? s : s - _2q[i];
        }
}

/* kern.cu:1038-1058 */
void orc_mont_sub(const int64_t *a, ptrdiff_t as, const int64_t *b, ptrdiff_t bs, int64_t *c, ptrdiff_t cs,
                  int C, int N, const int64_t *_2q)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < N; ++j) {
            const int64_t s = a[i * as + j] + _2q[i] - b[i * bs + j];
            c[i * cs + j] = (s < _2q[i]) ? s : s - _2q[i];
        }
}

/*
 * Forward negacyclic NTT, Cooley-Tukey, natural in -> bit-reversed out.
 * Stage loop = kern.cu:318-322 (one launch per stage); butterfly = kern.cu:257-274;
 * pair/twiddle indexing = cctx.py:89-113 (paint_butterfly_forward):
 *   stage logm: m = 2^logm, t = N/(2m); block i in [0,m): pairs (j, j+t), j in [2it, 2it+t),
 *   twiddle psi_rev[m+i].
 * psi: compact per-limb table [C][N] (row stride ps) of psi^bitrev(i) in Montgomery form,
 * i.e. the values the reference keeps "painted" in psi[C][logN][N/2] (cctx.py:336-338).
 */
void orc_ntt(int64_t *a, ptrdiff_t as, int C, int logN, const int64_t *psi, ptrdiff_t ps,
             const int64_t *_2q, const int64_t *ql, const int64_t *qh, const int64_t *kl, const int64_t *kh)
{
    const int N = 1 << logN;
    const int half = N >> 1;
#pragma omp parallel
    for (int logm = 0; logm < logN; ++logm) {
        const int logt = logN - 1 - logm;
        const int t = 1 << logt;
        const int m = 1 << logm;
#pragma omp for collapse(2) schedule(static)
        for (int i = 0; i < C; ++i)
            for (int b = 0; b < half; ++b) {
                const int blk = b >> logt, off = b & (t - 1);
                const int j = (blk << (logt + 1)) + off;
                int64_t *row = a + i * as;
                const int64_t U = row[j], O = row[j + t];
                const int64_t S = psi[i * ps + m + blk];
                const int64_t V = mm(S, O, ql[i], qh[i], kl[i], kh[i]);
                const int64_t up = U + V, um = U + _2q[i] - V;
                row[j] = (up < _2q[i]) ? up : up - _2q[i];
                row[j + t] = (um < _2q[i]) ? um : um - _2q[i];
            }
    }
}

/*
 * Inverse negacyclic NTT stages, Gentleman-Sande, bit-reversed in -> natural out, WITHOUT the
 * N^-1 scaling.  Stage loop kern.cu:521-525, butterfly kern.cu:454-472, indexing
 * cctx.py:116-142 (paint_butterfly_backward): level = 0..logN-1, t = 2^level, h = N/(2t);
 * block i in [0,h): pairs (j, j+t), j in [2it, 2it+t), twiddle ipsi_rev[h+i].
 */
static void intt_stages(int64_t *a, ptrdiff_t as, int C, int logN, const int64_t *ipsi, ptrdiff_t ps,
                        const int64_t *_2q, const int64_t *ql, const int64_t *qh, const int64_t *kl,
                        const int64_t *kh)
{
    const int N = 1 << logN;
    const int half = N >> 1;
#pragma omp parallel
    for (int level = 0; level < logN; ++level) {
        const int t = 1 << level;
        const int h = N >> (level + 1);
#pragma omp for collapse(2) schedule(static)
        for (int i = 0; i < C; ++i)
            for (int b = 0; b < half; ++b) {
                const int blk = b >> level, off = b & (t - 1);
                const int j = (blk << (level + 1)) + off;
                int64_t *row = a + i * as;
                const int64_t U = row[j], V = row[j + t];
                const int64_t S = ipsi[i * ps + h + blk];
                const int64_t um = U + _2q[i] - V;
                const int64_t O = (um < _2q[i]) ? um : um - _2q[i];
                row[j + t] = mm(S, O, ql[i], qh[i], kl[i], kh[i]);
                const int64_t up = U + V;
                row[j] = (up < _2q[i]) ? up : up - _2q[i];
            }
    }
}

/* exit_mode: 0 = intt (kern.cu:476-548), 1 = intt_exit (+redc, kern.cu:709-766),
 * 2 = intt_exit_reduce (+reduce, kern.cu:772-832), 3 = intt_exit_reduce_signed (+make_signed,
 * kern.cu:838-902). */
void orc_intt(int64_t *a, ptrdiff_t as, int C, int logN, const int64_t *ipsi, ptrdiff_t ps,
              const int64_t *Ninv, const int64_t *_2q, const int64_t *ql, const int64_t *qh,
              const int64_t *kl, const int64_t *kh, int exit_mode)
{
    const int N = 1 << logN;
    intt_stages(a, as, C, logN, ipsi, ps, _2q, ql, qh, kl, kh);
    orc_mont_enter(a, as, Ninv, C, N, ql, qh, kl, kh);
    if (exit_mode >= 1) orc_mont_redc(a, as, C, N, ql, qh, kl, kh);
    if (exit_mode >= 2) orc_reduce_2q(a, as, C, N, _2q);
    if (exit_mode >= 3) orc_make_signed(a, as, C, N, _2q);
}

/* enter_ntt: kern.cu:349-423 = mont_enter(Rs) then the forward stages. */
void orc_enter_ntt(int64_t *a, ptrdiff_t as, const int64_t *Rs, int C, int logN, const int64_t *psi,
                   ptrdiff_t ps, const int64_t *_2q, const int64_t *ql, const int64_t *qh,
                   const int64_t *kl, const int64_t *kh)
{
    orc_mont_enter(a, as, Rs, C, 1 << logN, ql, qh, kl, kh);
    orc_ntt(a, as, C, logN, psi, ps, _2q, ql, qh, kl, kh);
}

/* the bench's reference arm sets its own thread count: torchrun exports OMP_NUM_THREADS=1 to every rank */
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
