// graal_b200 -- B200-native GRAAL MCMC scoring path (sm_100a).  See include/graal_b200.h for the ABI
// and DESIGN.md for the formulation.  Compile with -fmad=false: the float32 geometry (mid-points,
// distances, norm factors) must round exactly like the reference's separate mul/add sequence;
// fused operations are written explicitly (fma()) where they are wanted.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <cuda/std/functional>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include <map>
#include <algorithm>

#include "../../include/graal_b200.h"
#include "moves.cuh"

#define GRAAL_BAND_RESYNC 256      // full recomputation of the cached band total every so many evaluations
#define GRAAL_VERSION "graal_b200 0.1 (sm_100a)"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int set_err(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return code;
}
#define CUDA_OK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) \
    return set_err(-2, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); } while (0)
#define CHECK_LAUNCH(ctx) do { (ctx)->launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) \
    return set_err(-3, "%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); } while (0)

// ------------------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------------------
// nd x nd tables over the distinct accu values (a_i, a_j) of the level, built on the host with the
// same IEEE float32 operations the reference performs per pixel:
//   t_norm[i*nd+j] = f32(f32(a_i*a_j) / nfpb)            (kernels3.cu:3065)
//   t_g   [i*nd+j] = f32(v_inter * t_norm)               the value of every clamped / trans pixel
//   t_logg[i*nd+j] = log((double) t_g)                   (-inf when t_g == 0: the pixel contributes 0)
struct Params { float kuhn, lm, c1, slope, d, d_max, fact, v_inter, nfpb; int nd;
                const float* t_norm; const float* t_g; const double* t_logg;
                // log-space evaluation (math mode 1): ln(c1*fact), ln(v_inter), slope as double, ln(t_norm)
                int mode; double ln_cf, ln_v, slope_d; const double* t_lnnorm;
                const double2* t_log; const double* t_exp;
                // tabulated law (math mode 2): piecewise quadratics {double a0; float a1, a2} of ln f(s) and of f(s),
                // one interval per float bit pattern with LAW_M mantissa bits, s in [2^LAW_EMIN, 2^LAW_EMAX) kb;
                // v_clamp = v_inter as double (the clamp of f), t_normd = t_norm as double
                const int4* t_lnf; const int4* t_f; double v_clamp; const double* t_normd;
                // uniform-accu levels with a monotone law (windowed contact pass): ln f(s) + ln norm - log g per interval, the
                // float bit patterns [LAW_SMIN_BITS, LAW_SMIN_BITS + fu_span) it is valid on (below d_max, inside the table,
                // above the clamp) and the in-band zone beyond the table [fu_zlo, fu_zlo + fu_zspan); fu_ok = 0: not available
                const int4* t_lnfu; unsigned fu_span, fu_zlo, fu_zspan; int fu_ok;
                // the same relative law on LAW7_OCT octaves below the end of the fast range, 2^LAW7_M intervals per octave
                // (32 KB: staged in shared memory by the windowed full pass); fu7_smin = bit pattern of its first distance
                const int4* t_lnfu7; unsigned fu7_smin, fu7_span;
                const int4* t_fu;     // same levels: f(s) * norm - g per interval (band excess of an in-band pair)
                };
#define LAW_M 9
#define LAW_EMIN (-12)
#define LAW_EMAX 11
#define LAW_NODES (((LAW_EMAX) - (LAW_EMIN)) << LAW_M)
#define LAW7_M 7
#define LAW7_OCT 16
#define LAW7_NODES (LAW7_OCT << LAW7_M)

// per-sub-frag geometry record of one slot (16 B, one 128-bit load):
//   mid  : mid-point in kb (float32, reference op order kernels3.cu:2997-3060)
//   id_c : contig id
//   stot : contig length in kb (float32 of l_cont_bp / 1000) -- used by circular contigs only
//   pk   : index of accu_true[0:8) | index of accu_quirk[8:16) (into the level's list of distinct accu
//          values) | local sub index[26:28) | circ[28]
struct __align__(16) Geo { float mid; int id_c; float stot; unsigned pk; };
#define MAX_ACCU_VALUES 256
// Candidates scored as a second-level delta against their predecessor (3 vs 2, 5 vs 4, 7 vs 6: the same insertion
// with the other orientation of fA -- only fA's own records differ): bit k set = candidate k is paired with k - 1.
#define PAIR_MASK 0xA8u
#define N_BAND_ROWS 16        // 13 candidates + 3 rows "partner of a paired candidate, restricted to the differing records"
__device__ __forceinline__ int pk_true(unsigned pk) { return pk & 255; }
__device__ __forceinline__ int pk_quirk(unsigned pk) { return (pk >> 8) & 255; }
__device__ __forceinline__ int pk_local(unsigned pk) { return (pk >> 26) & 3; }
__device__ __forceinline__ int pk_circ(unsigned pk) { return (pk >> 28) & 1; }
__device__ __forceinline__ bool geo_eq(const Geo& a, const Geo& b) {
    return __float_as_int(a.mid) == __float_as_int(b.mid) && a.id_c == b.id_c &&
           __float_as_int(a.stot) == __float_as_int(b.stot) && a.pk == b.pk;
}
__device__ __forceinline__ Geo ld_geo(const Geo* p) {
    int4 v = __ldg(reinterpret_cast<const int4*>(p));
    Geo g; g.mid = __int_as_float(v.x); g.id_c = v.y; g.stot = __int_as_float(v.z); g.pk = (unsigned)v.w;
    return g;
}

// rippe_contacts (kernels3.cu:120-133)
__device__ __forceinline__ float rippe_contacts(float s, const Params& p) {
    float r = 0.0f;
    if (s > 0.0f && s < p.d_max) {
        float x = s * p.lm / p.kuhn;
        r = (p.c1 * powf(s, p.slope) * expf((p.d - 2.0f) / (x * x + p.d))) * p.fact;      // pow(x, 2) == x*x exactly
    }
    return fmaxf(r, p.v_inter);
}

// rippe_contacts_circ (kernels3.cu:135-166)
__device__ __noinline__ float rippe_contacts_circ(float s, float s_tot, const Params& p) {
    float result = 0.0f;
    if (s > 0.0f && s < p.d_max) {
        float K = p.lm / p.kuhn;
        float nmax = K * 1.0f;
        float n = K * s * (s_tot - s) / s_tot;
        float norm_lin = rippe_contacts(s, p);
        float k3 = powf(p.kuhn, -3.0f);
        float norm_circ = (k3 * powf(nmax, p.slope) * expf((p.d - 2.0f) / (nmax * nmax + p.d))) * p.fact;
        float val = (k3 * powf(n, p.slope) * expf((p.d - 2.0f) / (n * n + p.d))) * p.fact;
        result = val * norm_lin / norm_circ;
    }
    return fmaxf(result, p.v_inter);
}

__device__ __forceinline__ double log_fact_term(float obf);

// ---- math mode 1: log-space float64 evaluation of the in-band expected value ------------------------
// ln(x) of a positive normal float32: 128-entry table of (1/c, ln c) on the top mantissa bits plus a
// degree-5 log1p series in r = m/c - 1, |r| <= 2^-8 (truncation < 1e-15).
__device__ __forceinline__ double fast_log_f32(float x, const double2* __restrict__ tab) {
    const unsigned b = __float_as_uint(x);
    const int e = (int)(b >> 23) - 127;
    const int i = (b >> 16) & 127;
    const double m = (double)__uint_as_float((b & 0x007fffffu) | 0x3f800000u);
    const double2 t = __ldg(&tab[i]);
    const double r = fma(m, t.x, -1.0);
    double q = fma(r, 0.2, -0.25);
    q = fma(q, r, 1.0 / 3.0);
    q = fma(q, r, -0.5);
    q = fma(q, r, 1.0);
    return fma((double)e, 0.6931471805599453094, fma(q, r, t.y));
}
// exp(t), |t| < 700: k = rint(t * 32/ln2), table of 2^(j/32), degree-4 series in the remainder
// (|r| <= ln2/64, truncation < 2e-12 relative).
__device__ __forceinline__ double fast_exp(double t, const double* __restrict__ tab) {
    const double kf = rint(t * 46.166241308446828384);
    const int k = (int)kf;
    double r = fma(kf, -0.021660849392498290, t);            // ln2/32 (hi)
    double q = fma(r, 1.0 / 24.0, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = fma(q, r, 1.0);
    const double base = __ldg(&tab[k & 31]) * q;
    const int sh = k >> 5;
    return __longlong_as_double(__double_as_longlong(base) + ((long long)sh << 52));
}
// Piecewise-quadratic tabulated function of s (math mode 2).  The interval index and the local abscissa come
// straight from the float bit pattern: interval = exponent | top LAW_M mantissa bits, and the remaining
// 23-LAW_M mantissa bits, moved to the top of a float32 mantissa with exponent 0, ARE the abscissa
// u = 1 + t in [1, 2) -- no conversion, no division.  One 16-byte entry per interval {double a0; float a1, a2}:
// value = a0 + u * (a1 + u * a2), the quadratic through the three Chebyshev nodes of the interval; the
// u-dependent part is below 1 % of a0, so float32 is enough for it.  One gather per evaluation.
// Relative error for the Rippe law: < 2e-9 (measured, profiles/README.md), far below one float32 ulp.
__device__ __forceinline__ bool law_in_table(float s) {
    const int e = (int)(__float_as_uint(s) >> 23) - 127;
    return e >= LAW_EMIN && e < LAW_EMAX;
}
__device__ __forceinline__ double law_interp(float s, const int4* __restrict__ tab) {
    const unsigned b = __float_as_uint(s);
    const int iv = (int)(b >> (23 - LAW_M)) - ((127 + LAW_EMIN) << LAW_M);
    const float u = __uint_as_float(0x3f800000u | ((b & ((1u << (23 - LAW_M)) - 1u)) << LAW_M));
    const int4 e = __ldg(&tab[iv]);
    return __hiloint2double(e.y, e.x) + (double)(u * fmaf(u, __int_as_float(e.w), __int_as_float(e.z)));
}

// ln of rippe_contacts(s) for 0 < s < d_max on a LINEAR contig, clamp included:
// max(ln(c1*fact) + slope*ln(s) + (d-2)/(x^2+d), ln v_inter) with x and the quotient in float32 exactly
// as the reference computes the argument of expf (kernels3.cu:126).
__device__ __forceinline__ double ln_rippe_inband(float s, const Params& p) {
    const float x = s * p.lm / p.kuhn;
    const float e = (p.d - 2.0f) / (x * x + p.d);
    const double lr = fma(p.slope_d, fast_log_f32(s, p.t_log), p.ln_cf + (double)e);
    return fmax(lr, p.ln_v);            // fmax drops a NaN / -inf ln_v (v_inter <= 0): max(r, v) = r
}

// expected contacts of the sub-frag pair (a, b); a belongs to the LOWER data bin (quirk Q1 side).
__device__ __forceinline__ float expected_pair(const Geo& a, const Geo& b, const Params& p) {
    if (a.id_c == b.id_c) {
        float s = fabsf(b.mid - a.mid);
        float norm = __ldg(&p.t_norm[pk_true(a.pk) * p.nd + pk_true(b.pk)]);
        float r = pk_circ(a.pk) ? rippe_contacts_circ(s, a.stot, p) : rippe_contacts(s, p);
        return r * norm;
    }
    return __ldg(&p.t_g[pk_quirk(a.pk) * p.nd + pk_true(b.pk)]);
}
// ex - g of an in-band cis pair (0 < s < d_max checked by the caller), as float64
__device__ __forceinline__ double band_excess(const Geo& a, const Geo& b, float s, const Params& p) {
    const int idx = pk_true(a.pk) * p.nd + pk_true(b.pk);
    const double g = (double)__ldg(&p.t_g[idx]);
    if (p.mode == 0 || pk_circ(a.pk)) {
        const float norm = __ldg(&p.t_norm[idx]);
        const float r = pk_circ(a.pk) ? rippe_contacts_circ(s, a.stot, p) : rippe_contacts(s, p);
        return (double)(r * norm) - g;
    }
    if (p.mode == 2 && law_in_table(s)) {
        const double f = law_interp(s, p.t_f);
        if (!(f > p.v_clamp)) return 0.0;                   // clamped: exactly the clamp value
        return f * __ldg(&p.t_normd[idx]) - g;
    }
    const double lr = ln_rippe_inband(s, p);
    if (!(lr > p.ln_v)) return 0.0;                         // clamped: exactly the clamp value
    return fast_exp(lr + __ldg(&p.t_lnnorm[idx]), p.t_exp) - g;
}
// ln(ex) of an in-band cis pair times ob (the contact term); falls back to the float32 chain for
// circular contigs and in math mode 0
__device__ __forceinline__ double inband_log_term(float s, float ob, float stot, int idx, int circ, const Params& p) {
    if (p.mode == 0 || circ) {
        const float norm = __ldg(&p.t_norm[idx]);
        const float r = circ ? rippe_contacts_circ(s, stot, p) : rippe_contacts(s, p);
        const float ex = r * norm;
        return (ex != 0.0f) ? (double)ob * log((double)ex) : log_fact_term(ob);
    }
    const double lr = ((p.mode == 2 && law_in_table(s)) ? fmax(law_interp(s, p.t_lnf), p.ln_v) : ln_rippe_inband(s, p))
                      + __ldg(&p.t_lnnorm[idx]);
    return (lr == lr && lr != -INFINITY) ? (double)ob * lr : ((lr != lr) ? lr : log_fact_term(ob));
}
// ob * ln(ex) of one stored contact between sub-frags a (row side = lower data bin) and b, or the
// lf(ob) correction when the expected value is 0 (kernels3.cu:197: such a pixel contributes 0)
// (scalar arguments: records passed by reference to an out-of-line function would be spilled to the stack on
//  the hot path of every caller)
__device__ __noinline__ double contact_log_term_general(float a_mid, int a_idc, float a_stot, unsigned a_pk,
                                                        float b_mid, int b_idc, unsigned b_pk, float ob, const Params& p) {
    const bool cis = a_idc == b_idc;
    const float s = fabsf(b_mid - a_mid);
    if (cis && s > 0.0f && s < p.d_max)
        return inband_log_term(s, ob, a_stot, pk_true(a_pk) * p.nd + pk_true(b_pk), pk_circ(a_pk), p);
    const double lg = __ldg(&p.t_logg[(cis ? pk_true(a_pk) : pk_quirk(a_pk)) * p.nd + pk_true(b_pk)]);
    return (lg != -INFINITY) ? (double)ob * lg : log_fact_term(ob);
}
// The same value with the common cases inline (tabulated law on a linear contig, finite clamp tables) and
// everything else behind ONE call: the 14 evaluations per contact of the delta pass stay small enough for
// the instruction cache.
__device__ __forceinline__ double contact_log_term(const Geo& a, const Geo& b, float ob, const Params& p) {
    const bool cis = a.id_c == b.id_c;
    const float s = fabsf(b.mid - a.mid);
    if (cis && s > 0.0f && s < p.d_max) {
        if (p.mode == 2 && !pk_circ(a.pk) && law_in_table(s)) {
            const double lr = fmax(law_interp(s, p.t_lnf), p.ln_v) + __ldg(&p.t_lnnorm[pk_true(a.pk) * p.nd + pk_true(b.pk)]);
            if (lr - lr == 0.0) return (double)ob * lr;                     // finite
        }
    } else {
        const double lg = __ldg(&p.t_logg[(cis ? pk_true(a.pk) : pk_quirk(a.pk)) * p.nd + pk_true(b.pk)]);
        if (lg - lg == 0.0) return (double)ob * lg;
    }
    return contact_log_term_general(a.mid, a.id_c, a.stot, a.pk, b.mid, b.id_c, b.pk, ob, p);
}
__device__ __noinline__ double inband_log_term_general(float s, float ob, float stot, int idx, int circ, const Params& p) {
    return inband_log_term(s, ob, stot, idx, circ, p);
}
// clamp value of the pair (true accus): what a cis pair beyond the band evaluates to
__device__ __forceinline__ float g_pair(const Geo& a, const Geo& b, const Params& p) {
    return __ldg(&p.t_g[pk_true(a.pk) * p.nd + pk_true(b.pk)]);
}

// the value every clamped / trans pixel takes: f32(v_inter * f32(f32(P)/nfpb))
__host__ __device__ __forceinline__ float g_clamp(int P, float v_inter, float nfpb) {
    float norm = (float)P / nfpb;
    return v_inter * norm;
}

// factorial (kernels3.cu:80-93) and the observation-only part of evaluate_likelihood_double (:198-203)
__device__ __forceinline__ float factorial_f32(float n) {
    float result = 1.0f;
    n = floorf(n);
    if (n < 10.0f) { for (int c = 1; c <= (int)n; c++) result = result * (float)c; }
    else result = powf(n, n) * expf(-n) * sqrtf((float)(2.0 * M_PI * (double)n));
    return result;
}
__device__ __forceinline__ double log_fact_term(float obf) {
    double ob = (double)obf;
    if (ob >= 15.0) return ob * log(ob) - ob + log(sqrt(ob * 2.0 * M_PI));
    if (ob > 0.0) return log((double)factorial_f32(obf));
    return 0.0;
}

// ------------------------------------------------------------------------------------------------
// reductions (deterministic two-stage: per-block partials, then one block sums them in order)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double block_sum(double v) {   // result valid in thread 0
    __shared__ double sh[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) sh[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
    if (w == 0) v = warp_sum(v);
    __syncthreads();
    return v;
}

// out[k] (op)= scale * sum_j partials[k * stride + j], j < n;  op: 0 set, 1 add
__global__ void k_reduce_partials(const double* __restrict__ partials, int n, int stride, double scale,
                                  double* __restrict__ out, int accumulate) {
    const int k = blockIdx.x;
    double v = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) v += partials[(size_t)k * stride + j];
    v = block_sum(v);
    if (threadIdx.x == 0) out[k] = (accumulate ? out[k] : 0.0) + scale * v;
}

// End of a proposal's delta, one launch for the 13 candidates: second stage of the three reductions and
//   band[k] = sum(cand band) - sum(base band);   out[k] = sum(contacts) - band[k];   candidate `copy_to` = candidate 0.
// A paired candidate (bit k of pair_mask) is composed from its partner k - 1:
//   band[k] = band[k-1] + (T_k - T'_k),  contacts[k] = contacts[k-1] + C_k
// with T_k / T'_k the band mass of the pairs touching a differing record in candidate k / in its partner (band
// rows k and 13 + (k-3)/2) and C_k the contact terms over the differing records.
__global__ void k_finish_delta(const double* __restrict__ p_contacts, int n_contacts, const double* __restrict__ p_cand, const double* __restrict__ p_base,
                               int n_band, int stride, double* __restrict__ out, double* __restrict__ band, int copy_to, unsigned pair_mask) {
    const int k = blockIdx.x;
    const bool paired = (pair_mask >> k) & 1u;
    const int src = paired ? k - 1 : ((copy_to > 0 && k == copy_to) ? 0 : k);      // the copied candidate sums candidate 0's partials itself
    double vc = 0.0, vn = 0.0, vb = 0.0;
    for (int j = threadIdx.x; j < n_contacts; j += blockDim.x) vc += p_contacts[(size_t)src * stride + j];
    for (int j = threadIdx.x; j < n_band; j += blockDim.x) { vn += p_cand[(size_t)src * stride + j]; vb += p_base[(size_t)src * stride + j]; }
    vc = block_sum(vc); vn = block_sum(vn); vb = block_sum(vb);
    double ck = 0.0, tk = 0.0, tp = 0.0;
    if (paired) {
        const int prow = GRAAL_N_CANDIDATES + (k - 3) / 2;
        for (int j = threadIdx.x; j < n_contacts; j += blockDim.x) ck += p_contacts[(size_t)k * stride + j];
        for (int j = threadIdx.x; j < n_band; j += blockDim.x) { tk += p_cand[(size_t)k * stride + j]; tp += p_cand[(size_t)prow * stride + j]; }
    }
    ck = block_sum(ck); tk = block_sum(tk); tp = block_sum(tp);
    if (threadIdx.x == 0) {
        double b = 0.0 + 1.0 * vn; b = b + (-1.0) * vb;
        double o = 0.0 + 1.0 * vc;
        if (paired) { b = b + (tk - tp); o = o + ck; }
        o = o + (-1.0) * b;
        band[k] = b; out[k] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// structure mutations
// ------------------------------------------------------------------------------------------------
__global__ void k_apply_move(const int* __restrict__ src, int* __restrict__ dst, int ld, int n,
                             int op, int fA, int fB, int aux, int max_id) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Bin s = load_bin(src, ld, i);
    bool written = true;
    switch (op) {
        case GRAAL_OP_COPY: break;
        case GRAAL_OP_FLIP: s = op_flip(s, i, fA); break;
        case GRAAL_OP_SWAP_ACTIV: s = op_swap_activity(s, i, fA, max_id); break;
        case GRAAL_OP_POP_OUT: { Bin P = load_bin(src, ld, fA); s = op_pop_out(s, i, P, max_id); } break;
        case GRAAL_OP_POP_IN_1: case GRAAL_OP_POP_IN_2: case GRAAL_OP_POP_IN_3: case GRAAL_OP_POP_IN_4: {
            Bin Pp = load_bin(src, ld, fA), Pi = load_bin(src, ld, fB);
            s = op_pop_in(op - GRAAL_OP_POP_IN_1 + 1, s, i, Pp, Pi, fA, fB, max_id, aux);
        } break;
        case GRAAL_OP_SPLIT: { Bin Pc = load_bin(src, ld, fA); s = op_split(s, i, Pc, aux, max_id); } break;
        case GRAAL_OP_PASTE: {
            Bin PA = load_bin(src, ld, fA), PB = load_bin(src, ld, fB);
            written = op_paste(s, i, PA, PB, fA, fB);
        } break;
    }
    if (written) store_bin(dst, ld, i, s);
}

__global__ void k_max_field(const int* __restrict__ v, int n, int* __restrict__ out) {
    int m = INT_MIN;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, v[i]);
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}
__global__ void k_set_int(int* p, int v) { *p = v; }
__global__ void k_set_double(double* p, double v) { *p = v; }
__global__ void k_init_stats(unsigned long long* a) { a[0] = 0ull; a[1] = 0ull; a[2] = (unsigned long long)(unsigned)INT_MAX; a[3] = (unsigned long long)(unsigned)INT_MIN; }

// the 13 candidates of new_perform_modificationS (cuda_lib_gl.py:841-954) in one pass.
__device__ __forceinline__ void build_candidates_body(const int* __restrict__ src, int* __restrict__ dst0, size_t slot_stride,
                                                      int ld, int n, int fA, int fB, const int* __restrict__ d_max_id,
                                                      int max_id_host, unsigned mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int M = (max_id_host >= 0) ? max_id_host : *d_max_id;
    const Bin C = load_bin(src, ld, i);
    const Bin CA = load_bin(src, ld, fA);
    const Bin CB = load_bin(src, ld, fB);
    if (mask & 0x1FDu) {   // anything that needs the popped-out structure (modes 0, 2..8)
        const Bin P = op_pop_out(C, i, CA, M);
        const Bin PA = op_pop_out(CA, fA, CA, M);
        const Bin PB = op_pop_out(CB, fB, CA, M);
        const int M2 = M + pop_out_new_ids(CA);
        if (mask & 1u) store_bin(dst0, ld, i, P);
        #pragma unroll
        for (int m = 2; m < 8; m++) {
            if (mask & (1u << m)) {
                Bin r = op_pop_in(1 + (m - 2) / 2, P, i, PA, PB, fA, fB, M2, (m & 1) ? -1 : 1);
                store_bin(dst0 + (size_t)m * slot_stride, ld, i, r);
            }
        }
        if (mask & (1u << 8)) store_bin(dst0 + 8 * slot_stride, ld, i, op_swap_activity(P, i, fA, M2));
    }
    if (mask & 2u) store_bin(dst0 + slot_stride, ld, i, op_flip(C, i, fA));
    if (mask & 0x1E00u) {
        #pragma unroll
        for (int uA = 0; uA < 2; uA++) {
            const Bin T1 = op_split(C, i, CA, uA, M);
            const Bin T1A = op_split(CA, fA, CA, uA, M);
            const Bin T1B = op_split(CB, fB, CA, uA, M);
            const int M1 = M + split_new_ids(CA, uA);
            #pragma unroll
            for (int uB = 0; uB < 2; uB++) {
                const int m = 9 + 2 * uA + uB;
                if (!(mask & (1u << m))) continue;
                Bin T2 = op_split(T1, i, T1B, uB, M1);
                const Bin T2A = op_split(T1A, fA, T1B, uB, M1);
                const Bin T2B = op_split(T1B, fB, T1B, uB, M1);
                if (op_paste(T2, i, T2A, T2B, fA, fB)) store_bin(dst0 + (size_t)m * slot_stride, ld, i, T2);
            }
        }
    }
}
__global__ void k_build_candidates(const int* __restrict__ src, int* __restrict__ dst0, size_t slot_stride,
                                   int ld, int n, int fA, int fB, const int* __restrict__ d_max_id,
                                   int max_id_host, unsigned mask) {
    build_candidates_body(src, dst0, slot_stride, ld, n, fA, fB, d_max_id, max_id_host, mask);
}
// the candidate drawn ON THE DEVICE (k_draw_candidates): sel = {sample_out, n_ok, status, id_f_sampled, op}; rebuilds candidate
// `op` of (fA, id_f_sampled) like test_copy_struct (cuda_lib_gl.py:1156-1183).  status != 0: no draw, nothing is built.
__global__ void k_build_candidates_sel(const int* __restrict__ src, int* __restrict__ dst0, size_t slot_stride,
                                       int ld, int n, int fA, const int* __restrict__ sel, const int* __restrict__ d_max_id, int max_id_host) {
    if (sel[2] != 0) return;
    const int op = sel[4];
    build_candidates_body(src, dst0, slot_stride, ld, n, fA, sel[3], d_max_id, max_id_host, op < 9 ? (1u << op) : 0x1E00u);
}
// commit of the drawn candidate: slot (first candidate slot + op) -> the current slot
__global__ void k_commit_sel(const int* __restrict__ cand0, size_t slot_stride, int* __restrict__ dst, int ld, int n, const int* __restrict__ sel) {
    if (sel[2] != 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    store_bin(dst, ld, i, load_bin(cand0 + (size_t)sel[4] * slot_stride, ld, i));
}
__global__ void k_add_selected_sel(double* total, const double* band_hist, const int* __restrict__ sel) {
    if (sel[2] != 0) return;
    *total += band_hist[sel[0]];          // band_hist[proposal][candidate] flat == sample_out
}

// ------------------------------------------------------------------------------------------------
// contig relabel
// ------------------------------------------------------------------------------------------------
__global__ void k_fill_int(int* p, int n, int v) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_first_index(const int* __restrict__ id_c, int n, int cap, int* __restrict__ first, int* __restrict__ err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = id_c[i];
    if (c < 0 || c >= cap) { atomicExch(err, 1); return; }
    // neighbouring bins mostly share their contig: one atomic per distinct id in the warp (its lowest lane = lowest bin)
    const unsigned peers = __match_any_sync(__activemask(), c);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicMin(&first[c], i);
}
// sort key of contig c = its length (sentinel: id not in use), value = c.  The radix sort is stable and the
// input is in id order, so the result is ordered by (length, old id) -- the reference's
// np.unique + argsort(kind='mergesort') order (cuda_lib_gl.py:1695-1749) -- with log2(n) key bits only.
__global__ void k_relabel_keys(const int* __restrict__ first, const int* __restrict__ l_cont, int cap, unsigned sentinel,
                               unsigned* __restrict__ keys, unsigned* __restrict__ vals) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cap) return;
    const int f = first[c];
    keys[c] = (f == INT_MAX) ? sentinel : min((unsigned)l_cont[f], sentinel - 1u);
    vals[c] = (unsigned)c;
}
__global__ void k_relabel_map(const unsigned* __restrict__ keys, const unsigned* __restrict__ vals, int cap, unsigned sentinel,
                              int* __restrict__ map, int* __restrict__ n_contigs) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= cap) return;
    if (keys[r] == sentinel) return;
    map[vals[r]] = r;
    if (r + 1 == cap || keys[r + 1] == sentinel) *n_contigs = r + 1;
}
__global__ void k_relabel_apply(int* __restrict__ id_c, int n, int cap, const int* __restrict__ map,
                                const int* __restrict__ n_contigs, int* __restrict__ max_id_a, int* __restrict__ max_id_b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { const int m = *n_contigs - 1; *max_id_a = m; if (max_id_b) *max_id_b = m; }
    if (i >= n) return;
    const int c = id_c[i];
    if (c >= 0 && c < cap) id_c[i] = map[c];            // an id outside [0, cap) was flagged by k_first_index (graal_sync reports it): left as it is
}

// ------------------------------------------------------------------------------------------------
// Fused step prologue (statistics + relabel) for genomes of at most RL_MAX contigs with ids below RL_RANGE: three
// launches instead of sixteen.  The per-step chain statistics -> relabel is a latency chain of tiny kernels on the
// critical path of every step (nothing else can start before the contig ids are final).
//   k_prologue_scan : per bin, the statistics of graal_state_stats and the first bin of every contig (first[] is kept
//                     at INT_MAX between calls: k_prologue_rank cleans what it reads)
//   k_prologue_rank : ONE block: the contigs in use compacted to (length, id) keys, sorted in shared memory (bitonic;
//                     keys are unique, so the result is the (length, old id) order of the stable sort of
//                     graal_relabel_contigs), rank -> map, statistics finalised
//   k_relabel_apply : id_c <- map[id_c]
// ------------------------------------------------------------------------------------------------
#define RL_MAX 4096
#define RL_RANGE 8192
__global__ void k_prologue_scan(const int* __restrict__ slot, int ld, int n, int* __restrict__ first, unsigned long long* __restrict__ acc, int* __restrict__ err) {
    unsigned long long heads = 0, sum = 0; int mn = INT_MAX, mx = INT_MIN;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int lc = slot[F_L_CONT * ld + i];
        mn = min(mn, lc); mx = max(mx, lc);
        if (slot[F_START_BP * ld + i] == 0) { heads++; sum += (unsigned long long)(long long)slot[F_L_CONT_BP * ld + i]; }
        const int c = slot[F_ID_C * ld + i];
        if (c < 0 || c >= RL_RANGE) { atomicExch(err, 1); continue; }
        const unsigned peers = __match_any_sync(__activemask(), c);
        if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicMin(&first[c], i);
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        heads += __shfl_down_sync(0xffffffffu, heads, o); sum += __shfl_down_sync(0xffffffffu, sum, o);
        mn = min(mn, __shfl_down_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (heads) { atomicAdd(&acc[0], heads); atomicAdd(&acc[1], sum); }
        atomicMin((int*)&acc[2], mn); atomicMax((int*)&acc[3], mx);
    }
}
__global__ void __launch_bounds__(1024)
k_prologue_rank(int* __restrict__ first, const int* __restrict__ l_cont, int range, int* __restrict__ map, int* __restrict__ d_ints,
                unsigned long long* __restrict__ acc, double* __restrict__ stats_out, int* __restrict__ max_id_out) {
    __shared__ unsigned long long key[RL_MAX];
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    for (int c = threadIdx.x; c < range; c += blockDim.x) {
        const int f = first[c];
        if (f != INT_MAX) {
            const int p = atomicAdd(&cnt, 1);
            if (p < RL_MAX) key[p] = ((unsigned long long)(unsigned)l_cont[f] << 32) | (unsigned)c;
            first[c] = INT_MAX;
        }
    }
    __syncthreads();
    int k = cnt;
    if (k > RL_MAX) { if (threadIdx.x == 0) atomicExch(&d_ints[2], 2); k = RL_MAX; }      // more contigs than the fused path holds: flagged
    int m = 1; while (m < k) m <<= 1;
    for (int i = k + threadIdx.x; i < m; i += blockDim.x) key[i] = ~0ull;
    __syncthreads();
    for (int size = 2; size <= m; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < (m >> 1); i += blockDim.x) {
                const int lo = ((i / stride) * stride << 1) + (i % stride), hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const unsigned long long a = key[lo], b = key[hi];
                if ((a > b) == up) { key[lo] = b; key[hi] = a; }
            }
            __syncthreads();
        }
    for (int r = threadIdx.x; r < k; r += blockDim.x) map[(int)(key[r] & 0xffffffffull)] = r;
    if (threadIdx.x == 0) {
        d_ints[1] = k; d_ints[0] = k - 1;
        if (max_id_out) *max_id_out = k - 1;
        if (stats_out) {
            stats_out[0] = (double)k;
            stats_out[1] = (double)*(const int*)&acc[2];
            stats_out[2] = acc[0] ? (double)acc[1] / (double)acc[0] : 0.0;
            stats_out[3] = (double)*(const int*)&acc[3];
        }
        acc[0] = 0ull; acc[1] = 0ull; acc[2] = (unsigned long long)(unsigned)INT_MAX; acc[3] = (unsigned long long)(unsigned)INT_MIN;
    }
}

// ------------------------------------------------------------------------------------------------
// geometry: per-sub-frag records of a slot (unique bins: frag index == data bin index)
// ------------------------------------------------------------------------------------------------
struct LevelView {
    const int4* sub_id; const float* sub_len; const int* sub_accu;   // [N] int4, [N*3], [N*3]
    const unsigned char* accu_idx;                                  // [N*3] index of sub_accu in the distinct-value list
    const unsigned char* dup;                                       // [N] 1: the data bin has repeat copies (nullptr: none)
    int n_data;                                                     // N: frags >= N are repeat copies
};
// Bins handled by the sparse (unique x unique) kernels; every pixel touching a duplicated data bin goes
// through the per-pixel repeat path (k_repeat_pixels).
__device__ __forceinline__ bool eligible(const LevelView& lv, int f) {
    return lv.dup == nullptr || (f < lv.n_data && !lv.dup[f]);
}
#define PK_EXCLUDED (1u << 29)

__device__ __forceinline__ void bin_geometry(const int* __restrict__ slot, int ld, int bin, const LevelView& lv,
                                             Geo* __restrict__ geo, unsigned short* __restrict__ cid16 = nullptr,
                                             float* __restrict__ mid32 = nullptr, int2* __restrict__ cm = nullptr) {
    const int id_d = slot[F_ID_D * ld + bin];
    const bool elig = eligible(lv, bin);
    if (!elig && bin >= lv.n_data) return;                   // repeat copy: its data sub-frags belong to the original
    const int4 sid = lv.sub_id[id_d];
    const int lim = sid.w - 1;
    const float len[3] = { lv.sub_len[id_d * 3], lv.sub_len[id_d * 3 + 1], lv.sub_len[id_d * 3 + 2] };
    const int acc[3] = { lv.accu_idx[id_d * 3], lv.accu_idx[id_d * 3 + 1], lv.accu_idx[id_d * 3 + 2] };
    const int ori = slot[F_ORI * ld + bin];
    const float start_kb = __int2float_rn(slot[F_START_BP * ld + bin]) / 1000.0f;
    const int id_c = slot[F_ID_C * ld + bin];
    const unsigned circ = slot[F_CIRC * ld + bin] == 1 ? 1u : 0u;
    // contig length (kb): only rippe_contacts_circ reads it, so it is recorded for circular contigs only --
    // a linear contig whose length changed but whose bins did not move keeps bit-identical records
    const float stot = circ ? __int2float_rn(slot[F_L_CONT_BP * ld + bin]) / 1000.0f : 0.0f;
    const int acc_last = acc[lim];
    float accu = 0.0f;
    #pragma unroll
    for (int i = 0; i < 3; i++) {
        if (i > lim) break;
        const int loc = (ori == 1) ? i : lim - i;           // list order -> local sub index
        const float l = (loc == 0) ? len[0] : (loc == 1 ? len[1] : len[2]);
        float mid;
        if (i == 0) { mid = start_kb + l / 2.0f; accu = start_kb + l; }
        else { mid = accu + l / 2.0f; accu = accu + l; }
        const int at = (loc == 0) ? acc[0] : (loc == 1 ? acc[1] : acc[2]);
        const int aq = (ori == 1) ? at : acc_last;
        Geo g; g.mid = mid; g.id_c = id_c; g.stot = stot;
        g.pk = (unsigned)at | ((unsigned)aq << 8) | ((unsigned)loc << 26) | (circ << 28) | (elig ? 0u : PK_EXCLUDED);
        const int sub = (loc == 0) ? sid.x : (loc == 1 ? sid.y : sid.z);
        geo[sub] = g;
        if (cid16) { cid16[sub] = (unsigned short)min(id_c < 0 ? 65535 : id_c, 65535); mid32[sub] = mid; }   // classification tables
        if (cm) cm[sub] = make_int2(id_c, __float_as_int(mid));                                                // {contig id, mid-point}: one 8-byte gather
    }
}

__global__ void k_geometry_all(const int* __restrict__ slot, int ld, int n, LevelView lv, Geo* __restrict__ geo,
                               unsigned short* __restrict__ cid16, float* __restrict__ mid32, int2* __restrict__ cm) {
    const int bin = blockIdx.x * blockDim.x + threadIdx.x;
    if (bin < n) bin_geometry(slot, ld, bin, lv, geo, cid16, mid32, cm);
}

// ------------------------------------------------------------------------------------------------
// full likelihood, contact part: stream the row-segmented contact list once
// ------------------------------------------------------------------------------------------------
#define CHUNK 2048       // entries per warp-chunk
#define QCAP 64          // per-warp queue of in-band entries (warp compaction)

// One in-band cis entry waiting for the expensive evaluation (powf / expf / fp64 log).
struct __align__(16) Pending { float s; float ob; float stot; unsigned key; };   // key = norm index | circ << 31

__device__ __forceinline__ double inband_term(const Pending& q, const Params& p) {
    return inband_log_term(q.s, q.ob, q.stot, (int)(q.key & 0x7fffffffu), (int)(q.key >> 31), p);
}

// Streams the contact list once.  Every warp owns one contiguous span of entries (one row search),
// keeps UNROLL coalesced 256-B loads of the list plus the dependent 16-B record gathers in flight per
// lane, and classifies each entry: trans / clamped-cis entries (the vast majority) cost one table
// lookup (ob * log g); in-band cis entries are compacted into a per-warp shared-memory queue and
// evaluated 32 at a time, so the expensive path (powf / expf / fp64 log) always runs with full lanes.
#define UNROLL 4
#define FC_MIN_BLOCKS 4
__device__ __forceinline__ int2 ld_stream(const int2* p) {      // read-once data: evict-first
    int2 v; asm volatile("ld.global.cs.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p)); return v;
}
__global__ void __launch_bounds__(256, FC_MIN_BLOCKS)
k_full_contacts(const long long* __restrict__ rowptr, const int2* __restrict__ contacts, long long E, int W,
                const Geo* __restrict__ geo, const __grid_constant__ Params p, double* __restrict__ partials) {
    __shared__ Pending queue[8][QCAP];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    // contiguous span of this warp, in units of 32 * UNROLL entries
    const long long n_groups = (E + 32 * UNROLL - 1) / (32 * UNROLL);
    const long long g0 = n_groups * warp / n_warps, g1 = n_groups * (warp + 1) / n_warps;
    const long long e0 = g0 * 32 * UNROLL, e1 = min(E, g1 * 32 * UNROLL);
    Pending* q = queue[wib];
    int qn = 0;
    double acc = 0.0;
    if (e0 < e1) {
        int lo = 0, hi = W;      // first row whose segment ends after e0
        while (lo < hi) { int mid = (lo + hi) >> 1; if (__ldg(&rowptr[mid + 1]) > e0) hi = mid; else lo = mid + 1; }
        int row = lo;
        long long row_end = __ldg(&rowptr[row + 1]);
        Geo gr = ld_geo(&geo[row]);
        for (long long eb = e0; eb < e1; eb += 32 * UNROLL) {
            int2 ce[UNROLL]; Geo gc[UNROLL];
            #pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const long long e = eb + u * 32 + lane;
                ce[u] = (e < e1) ? ld_stream(&contacts[e]) : make_int2(0, 0);
            }
            #pragma unroll
            for (int u = 0; u < UNROLL; u++) gc[u] = ld_geo(&geo[ce[u].x]);
            #pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const long long e = eb + u * 32 + lane;
                bool inband = false;
                Pending mine;
                if (e < e1) {
                    if (e >= row_end) {
                        do { row++; row_end = __ldg(&rowptr[row + 1]); } while (e >= row_end);
                        gr = ld_geo(&geo[row]);
                    }
                    const float ob = __int_as_float(ce[u].y);
                    const bool cis = gr.id_c == gc[u].id_c;
                    const float s = fabsf(gc[u].mid - gr.mid);
                    const bool excl = ((gr.pk | gc[u].pk) & PK_EXCLUDED) != 0u;      // pixel of a duplicated bin: repeat path
                    inband = !excl && cis && s > 0.0f && s < p.d_max;
                    const int idx = (cis ? pk_true(gr.pk) : pk_quirk(gr.pk)) * p.nd + pk_true(gc[u].pk);
                    if (excl) {
                    } else if (inband) {
                        mine.s = s; mine.ob = ob; mine.stot = gr.stot; mine.key = (unsigned)idx | ((unsigned)pk_circ(gr.pk) << 31);
                    } else {
                        const double lg = __ldg(&p.t_logg[idx]);
                        acc += (lg != -INFINITY) ? (double)ob * lg : log_fact_term(ob);     // -inf marks g == 0
                    }
                }
                const unsigned ballot = __ballot_sync(0xffffffffu, inband);
                if (inband) q[qn + __popc(ballot & lt_mask)] = mine;
                qn += __popc(ballot);
                __syncwarp();
                if (qn >= 32) {
                    qn -= 32;
                    acc += inband_term(q[qn + lane], p);
                    __syncwarp();
                }
            }
        }
    }
    if (lane < qn) acc += inband_term(q[lane], p);
    acc = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// Uniform-accu variant (every sub-frag has the same accu, e.g. level 1): every trans / clamped entry
// has the SAME log g, so their total is log g * sum(ob) with sum(ob) a level constant; the kernel only
// CLASSIFIES entries and evaluates the in-band ones, accumulating ob * (ln ex - log g).
//  * the list is cut into groups of GROUP = 128 entries; warp w takes groups w, w + n_warps, ... so that at any
//    time the chip streams ONE contiguous region of the list (DRAM page locality), 4 coalesced 256-B
//    loads in flight per warp, marked evict-first (measured: 4 loads x 32 warps / SM beats 8 x 24);
//  * the row of the first entry of every group is precomputed at bind time (no search);
//  * classification reads the 2-byte contig id of the partner (a W*2-byte table that lives in L1/L2)
//    and only cis entries -- which sit close to their row, i.e. share cache lines -- fetch the partner
//    mid-point; all gathers of a group are issued before the first one is used.
#define FC_UNROLL 4
#define GROUP (32 * FC_UNROLL)
__global__ void k_group_rows(const long long* __restrict__ rowptr, long long E, int W, int n_groups, int* __restrict__ group_row) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const long long e0 = (long long)g * GROUP;
    int lo = 0, hi = W;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (rowptr[mid + 1] > e0) hi = mid; else lo = mid + 1; }
    group_row[g] = lo;
}

#define FCS_THREADS 768
template <bool SMEM_CID>
__global__ void __launch_bounds__(SMEM_CID ? FCS_THREADS : 256, SMEM_CID ? 1 : 4)
k_full_contacts_uniform(const long long* __restrict__ rowptr, const int2* __restrict__ contacts, long long E, int W,
                        const int* __restrict__ group_row, int n_groups,
                        const Geo* __restrict__ geo, const unsigned short* __restrict__ cid16, const float* __restrict__ mid32,
                        const __grid_constant__ Params p, double lg, double* __restrict__ partials) {
    // SMEM_CID: the W x 2-byte contig-id table is staged in shared memory (one CTA per SM): the cis test
    // becomes a shared-memory gather (~3-way bank conflicts) instead of 32 L1 wavefronts per warp-load
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    __shared__ Pending queue_static[SMEM_CID ? 1 : 8][SMEM_CID ? 1 : QCAP];
    Pending* queue = SMEM_CID ? reinterpret_cast<Pending*>(smem_dyn) : &queue_static[0][0];
    const unsigned short* cid_tab = cid16;
    if (SMEM_CID) {
        unsigned short* cid_s = reinterpret_cast<unsigned short*>(smem_dyn + (size_t)(FCS_THREADS / 32) * QCAP * sizeof(Pending));
        const int n16 = (W + 7) / 8;
        const uint4* src = reinterpret_cast<const uint4*>(cid16);
        uint4* dst = reinterpret_cast<uint4*>(cid_s);
        for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(&src[i]);
        __syncthreads();
        cid_tab = cid_s;
    }
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    Pending* q = queue + wib * QCAP;
    int qn = 0;
    double acc = 0.0;
    for (int g = warp; g < n_groups; g += n_warps) {
        const long long e0 = (long long)g * GROUP;
        const int len = (int)min((long long)GROUP, E - e0);
        int2 ce[FC_UNROLL];
        #pragma unroll
        for (int u = 0; u < FC_UNROLL; u++) {
            const int r = u * 32 + lane;
            ce[u] = (r < len) ? ld_stream(&contacts[e0 + r]) : make_int2(0, 0);
        }
        int row = __ldg(&group_row[g]);
        int row_end_rel = (int)min(__ldg(&rowptr[row + 1]) - e0, (long long)INT_MAX);
        const Geo g0 = ld_geo(&geo[row]);
        float r_mid = g0.mid, r_stot = g0.stot; int r_idc = g0.id_c; unsigned r_circ = (unsigned)pk_circ(g0.pk) << 31;
        unsigned cc[FC_UNROLL];
        #pragma unroll
        for (int u = 0; u < FC_UNROLL; u++) cc[u] = SMEM_CID ? (unsigned)cid_tab[ce[u].x] : (unsigned)__ldg(&cid16[ce[u].x]);
        // per-lane row of each of its 8 entries (rows are ~100s of entries long: the cursor rarely moves)
        bool cis[FC_UNROLL]; float rm[FC_UNROLL], rs[FC_UNROLL]; unsigned rc[FC_UNROLL];
        #pragma unroll
        for (int u = 0; u < FC_UNROLL; u++) {
            const int r = u * 32 + lane;
            if (r < len && r >= row_end_rel) {
                long long re;
                do { row++; re = __ldg(&rowptr[row + 1]) - e0; } while ((long long)r >= re);
                row_end_rel = (int)min(re, (long long)INT_MAX);
                const Geo gg = ld_geo(&geo[row]);
                r_mid = gg.mid; r_idc = gg.id_c; r_stot = gg.stot; r_circ = (unsigned)pk_circ(gg.pk) << 31;
            }
            bool c = (int)cc[u] == r_idc;
            if (cc[u] == 65535u || (unsigned)r_idc >= 65535u) c = ld_geo(&geo[ce[u].x]).id_c == r_idc;   // ids beyond 16 bits: exact path
            cis[u] = c && r < len;
            rm[u] = r_mid; rs[u] = r_stot; rc[u] = r_circ;
        }
        float pm[FC_UNROLL];
        #pragma unroll
        for (int u = 0; u < FC_UNROLL; u++) pm[u] = cis[u] ? __ldg(&mid32[ce[u].x]) : 0.0f;
        #pragma unroll
        for (int u = 0; u < FC_UNROLL; u++) {
            const float s = fabsf(pm[u] - rm[u]);
            const bool inband = cis[u] && s > 0.0f && s < p.d_max;
            const unsigned ballot = __ballot_sync(0xffffffffu, inband);
            if (inband) { Pending m; m.s = s; m.ob = __int_as_float(ce[u].y); m.stot = rs[u]; m.key = rc[u]; q[qn + __popc(ballot & lt_mask)] = m; }
            qn += __popc(ballot);
            __syncwarp();
            if (qn >= 32) {
                qn -= 32;
                const Pending m = q[qn + lane];
                acc += inband_term(m, p) - (double)m.ob * lg;
                __syncwarp();
            }
        }
    }
    if (lane < qn) { const Pending m = q[lane]; acc += inband_term(m, p) - (double)m.ob * lg; }
    acc = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// Math mode 2 on a uniform-accu level: the in-band term is one table gather and a dozen instructions, cheaper
// than compacting the in-band lanes through shared memory -- every lane evaluates its own entries,
// branch-free; circular contigs and distances outside the table take the general path (rare).
//   acc += ob * (max(ln f(s), ln v) + ln norm - log g)   for in-band cis entries
// Per entry: the streamed 8-byte contact, ONE 8-byte gather {contig id, mid-point} of the partner, one 16-byte
// table gather.  Entries past the end of the list are loaded as (col 0, ob 0) and contribute exactly 0.
// The eight 256-byte stream loads of a group are staged in registers.  Measured slower on B200 (profiles/README.md):
// per-warp rings in shared memory filled by 2 KB cp.async.bulk copies + mbarrier (0.212 ms per pass: the per-SM
// bulk-copy rate is the limit) or by 16-byte cp.async four groups ahead (0.145 ms) against 0.123 ms here --
// the pass waits on its gathers and on instruction issue, not on the stream.
__global__ void __launch_bounds__(256, 4)
k_full_contacts_direct(const long long* __restrict__ rowptr, const int2* __restrict__ contacts, long long E,
                       const int* __restrict__ group_row, int n_groups,
                       const Geo* __restrict__ geo, const int2* __restrict__ cm,
                       const __grid_constant__ Params p, double lg, double* __restrict__ partials) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const double cst = __ldg(&p.t_lnnorm[0]) - lg;
    const bool fast_ok = cst - cst == 0.0;                      // finite
    const double ln_v = p.ln_v;
    const float d_max = p.d_max;
    const int4* __restrict__ t_lnf = p.t_lnf;
    double acc = 0.0;
    for (int g = warp; g < n_groups; g += n_warps) {
        const long long e0 = (long long)g * GROUP;
        const int len = (int)min((long long)GROUP, E - e0);
        const int2* __restrict__ cg = contacts + e0;
        int2 ce[FC_UNROLL];
        #pragma unroll
        for (int u = 0; u < FC_UNROLL; u++) {
            const int r = u * 32 + lane;
            ce[u] = (r < len) ? ld_stream(&cg[r]) : make_int2(0, 0);
        }
        int row = __ldg(&group_row[g]);
        int row_end_rel = (int)min(__ldg(&rowptr[row + 1]) - e0, (long long)GROUP);
        int2 rr = __ldg(&cm[row]);                                  // {contig id, mid-point} of the row
        int r_slow = pk_circ(__ldg(&geo[row].pk));
        int2 pc[FC_UNROLL];
        #pragma unroll
        for (int u = 0; u < FC_UNROLL; u++) pc[u] = __ldg(&cm[ce[u].x]);
        unsigned slow = 0u;
        #pragma unroll
        for (int u = 0; u < FC_UNROLL; u++) {
            const int r = u * 32 + lane;
            if (r >= row_end_rel && r < len) {                      // rows are ~100s of entries long: the cursor rarely moves
                long long re;
                do { row++; re = __ldg(&rowptr[row + 1]) - e0; } while ((long long)r >= re);
                row_end_rel = (int)min(re, (long long)GROUP);
                rr = __ldg(&cm[row]);
                r_slow = pk_circ(__ldg(&geo[row].pk));
            }
            const float s = fabsf(__int_as_float(pc[u].y) - __int_as_float(rr.y));
            const bool inband = pc[u].x == rr.x && s > 0.0f && s < d_max;
            const bool fast = inband && fast_ok && !r_slow && law_in_table(s);
            if (inband && !fast) slow |= 1u << u;
            const double lr = fmax(law_interp(fast ? s : 1.0f, t_lnf), ln_v) + cst;
            acc = fma((double)(fast ? __int_as_float(ce[u].y) : 0.0f), lr, acc);
        }
        if (slow) {                                                 // circular contig / outside the table / non-finite tables
            #pragma unroll
            for (int u = 0; u < FC_UNROLL; u++) {
                if (!((slow >> u) & 1u) || u * 32 + lane >= len) continue;
                const long long e = e0 + u * 32 + lane;             // the row of entry u is not kept per entry: find it again
                int lo = __ldg(&group_row[g]);
                while (__ldg(&rowptr[lo + 1]) <= e) lo++;
                const Geo gr = ld_geo(&geo[lo]);
                const float ob = __int_as_float(ce[u].y);
                acc += inband_log_term_general(fabsf(__int_as_float(pc[u].y) - gr.mid), ob, gr.stot, 0, pk_circ(gr.pk), p) - (double)ob * lg;
            }
        }
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// ------------------------------------------------------------------------------------------------
// Windowed contact pass (uniform-accu levels, math mode 2, monotone law): the default full-likelihood kernel.
//
// On a uniform-accu level every entry that is NOT an in-band cis pair contributes the level constant
// ob * log g, so the pass only has to find the in-band entries.  An entry (row r, column c) can be in band
// only if c lies inside the SUB-FRAG INDEX HULL of the bins within d_max of r's bin on r's contig in the
// current genome -- a per-row interval [wlo, wlo + wspan] computed once per pass from the position order
// (k_windows below).  The test is two integer instructions on the streamed column, no gather: far entries
// (trans pairs, cis pairs beyond the band: the bulk of a real list) cost the stream load and the test, and
// only entries inside the window take the exact path (partner {contig id, mid-point} gather, float32
// distance, table evaluation).  The window is conservative (a superset of the in-band partners, with a
// margin for the float32 rounding of the mid-points), so the result is the same sum as the gather-everything
// kernel; a scrambled genome only widens the windows.
//
// Work items are rows cut into chunks of <= ITEM_CHUNK entries (static, built at bind): one warp per item,
// items interleaved over the warps so that the chip streams one region of the list at a time; the row record
// {contig id, mid-point, window} of the next item is prefetched while the current one is processed.
// ------------------------------------------------------------------------------------------------
#define ITEM_CHUNK 2048
__global__ void k_item_count(const long long* __restrict__ rowptr, int W, int* __restrict__ cnt) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= W) return;
    const long long len = rowptr[r + 1] - rowptr[r];
    cnt[r] = (int)((len + ITEM_CHUNK - 1) / ITEM_CHUNK);
}
__global__ void k_item_fill(const long long* __restrict__ rowptr, int W, const int* __restrict__ off, int4* __restrict__ items) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= W) return;
    const long long e0 = rowptr[r], e1 = rowptr[r + 1];
    int k = off[r];
    for (long long e = e0; e < e1; e += ITEM_CHUNK, k++)
        items[k] = make_int4(r, (int)min((long long)ITEM_CHUNK, e1 - e), (int)(unsigned)(e & 0xffffffffll), (int)(e >> 32));
}

// Position order of a slot with the per-position fields the window pass scans: bin, start (bp), sub-frag index
// range of the bin, contig.  Unfilled positions (inconsistent state) keep the "everything" hull.
__global__ void k_order_init(int n, int W, int* __restrict__ order, int* __restrict__ o_start, int2* __restrict__ o_sub, int* __restrict__ o_cid, int* __restrict__ bad) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) *bad = 0;
    if (k >= n) return;
    order[k] = -1; o_start[k] = 0; o_sub[k] = make_int2(0, W - 1); o_cid[k] = -1;
}
__global__ void k_order_fill2(const int* __restrict__ slot, int ld, int n, int cap, const int* __restrict__ cont_off, LevelView lv,
                              int* __restrict__ order, int* __restrict__ o_start, int2* __restrict__ o_sub, int* __restrict__ o_cid, int* __restrict__ bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = slot[F_ID_C * ld + i];
    const int k = (c >= 0 && c < cap) ? cont_off[c] + slot[F_POS * ld + i] : -1;
    if (k < 0 || k >= n) { *bad = 1; return; }
    const int4 sid = lv.sub_id[slot[F_ID_D * ld + i]];
    order[k] = i; o_start[k] = slot[F_START_BP * ld + i]; o_sub[k] = make_int2(sid.x, sid.x + sid.w - 1); o_cid[k] = c;
}
// hull of the sub-frag index ranges of every aligned block of 32 order positions
__global__ void k_hull_blocks(const int2* __restrict__ o_sub, int n, int2* __restrict__ blk) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    int lo = INT_MAX, hi = -1;
    if (g < n) { const int2 s = o_sub[g]; lo = s.x; hi = s.y; }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((threadIdx.x & 31) == 0 && g < n) blk[g >> 5] = make_int2(lo, hi);
}
// Row records of the windowed pass, one per sub-frag: {contig id, mid-point bits, wlo, wspan | slow << 31}.
// One thread per order position g: the bins [lo, hi] of the same contig that can hold a sub-frag within d_max
// of a sub-frag of this bin (bin ends compared in bp with a margin of 0.01 % + 100 bp over the float32
// rounding of the mid-points), then the hull of their sub-frag index ranges from the block table.
__global__ void __launch_bounds__(256)
k_windows(const int* __restrict__ order, const int* __restrict__ o_start, const int2* __restrict__ o_sub, const int* __restrict__ o_cid,
          const int2* __restrict__ blk, const int* __restrict__ cont_off, const int* __restrict__ cont_len,
          const int* __restrict__ slot, int ld, int n, int W, long long dmax_bp, const int2* __restrict__ cm,
          int4* __restrict__ hdr, int* __restrict__ bad) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const int bin = order[g];
    if (bin < 0) { *bad = 1; return; }
    const int c = o_cid[g];
    const int cs = cont_off[c], ce = cs + cont_len[c];
    int wl = 0, wh = W - 1;
    if (cs <= g && g < ce && ce <= n) {
        const long long st = o_start[g], en = st + slot[F_LEN_BP * ld + bin];
        int a = cs, b = g;                      // lo: first position whose bin ends after st - dmax
        while (a < b) { const int m = (a + b) >> 1; if ((long long)o_start[m + 1] > st - dmax_bp) b = m; else a = m + 1; }
        const int lo = a;
        a = g; b = ce - 1;                      // hi: last position whose bin starts before en + dmax
        while (a < b) { const int m = (a + b + 1) >> 1; if ((long long)o_start[m] < en + dmax_bp) a = m; else b = m - 1; }
        const int hi = a;
        wl = INT_MAX; wh = -1;
        int p = lo;
        for (; p <= hi && (p & 31); p++) { const int2 s = o_sub[p]; wl = min(wl, s.x); wh = max(wh, s.y); }
        for (; p + 31 <= hi; p += 32) { const int2 s = blk[p >> 5]; wl = min(wl, s.x); wh = max(wh, s.y); }
        for (; p <= hi; p++) { const int2 s = o_sub[p]; wl = min(wl, s.x); wh = max(wh, s.y); }
    } else *bad = 1;
    const unsigned slow = slot[F_CIRC * ld + bin] == 1 ? 0x80000000u : 0u;
    const int2 s = o_sub[g];
    for (int sub = s.x; sub <= s.y; sub++) {
        const int2 r = cm[sub];
        hdr[sub] = make_int4(r.x, r.y, wl, (int)((unsigned)(wh - wl) | slow));
    }
}

// fast law of the windowed pass: ob * (ln f(s) + ln norm - log g) for SMIN <= s < s_hi from the table whose a0
// already holds the two constants; [zlo, zlo + zspan): in-band distances beyond the table (general path)
struct FastLaw { const int4* tab; unsigned smin_bits, span, zlo, zspan; };
#define LAW_SMIN_BITS ((unsigned)(127 + LAW_EMIN) << 23)

// per-item copy of the row records (item i -> record of its row): the pass then reads items[i] and ihdr[i] with
// no dependent load between them
__global__ void k_item_hdr(const int4* __restrict__ items, int n_items, const int4* __restrict__ hdr, int4* __restrict__ ihdr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_items) ihdr[i] = hdr[items[i].x];
}

// rare paths of the windowed pass, out of line: one in-band entry whose distance is outside the table ...
__device__ __noinline__ double win_slow_entry(int col, float ob, float r_mid, const int2* __restrict__ cm, const Params& p, double lg) {
    const float s = fabsf(__int_as_float(__ldg(&cm[col]).y) - r_mid);
    if (!(s > 0.0f && s < p.d_max)) return 0.0;
    return inband_log_term(s, ob, 0.0f, 0, 0, p) - (double)ob * lg;
}
// ... and a whole item whose row lies on a circular contig (float32 chain of rippe_contacts_circ)
__device__ __noinline__ double win_circular_item(const int2* __restrict__ cp, int len, int lane, int wlo, unsigned wspan, int r_id, float r_mid,
                                                 float r_stot, const int2* __restrict__ cm, const Params& p, double lg) {
    double acc = 0.0;
    for (int off = lane; off < len; off += 32) {
        const int2 ce = ld_stream(cp + off);
        if ((unsigned)(ce.x - wlo) > wspan) continue;
        const int2 pc = __ldg(&cm[ce.x]);
        const float s = fabsf(__int_as_float(pc.y) - r_mid);
        if (pc.x != r_id || !(s > 0.0f && s < p.d_max)) continue;
        const float ob = __int_as_float(ce.y);
        acc += inband_log_term(s, ob, r_stot, 0, 1, p) - (double)ob * lg;
    }
    return acc;
}

// UN: stream loads in flight per warp (32 entries each); SUB: the exact path runs on groups of SUB loads behind ONE
// uniform branch, straight-line inside -- the SUB partner gathers, then the SUB table gathers are in flight together
// (a branch per load would serialise the two dependent gathers of every load).
// TM: mantissa bits of the law table index.  LAW_M: the global table (L1 / L2 gathers); LAW7_M: the coarse table, staged
// in shared memory by every CTA (the table gather leaves the global load path; |error| of ln f <= 1e-8, oscillating)
template <int UN, int MINB, int SUB, int TM = LAW_M>
__global__ void __launch_bounds__(256, MINB)
k_full_contacts_win(const int4* __restrict__ items, const int4* __restrict__ ihdr, int n_items, const int2* __restrict__ contacts,
                    const int2* __restrict__ cm, const int* __restrict__ bad, int W,
                    const Geo* __restrict__ geo, const FastLaw fl, const __grid_constant__ Params p, double lg, double* __restrict__ partials) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const bool all_w = *bad != 0;                                   // inconsistent position order: every window is the whole level
    constexpr bool STAB = TM != LAW_M;
    __shared__ int4 stab[STAB ? LAW7_NODES : 1];
    if (STAB) {
        for (int k = threadIdx.x; k < LAW7_NODES; k += blockDim.x) stab[STAB ? k : 0] = __ldg(&fl.tab[k]);
        __syncthreads();
    }
    const int4* __restrict__ tab = fl.tab;
    const unsigned smin = fl.smin_bits, span = fl.span, zlo = fl.zlo, zspan = fl.zspan;
    double accd = 0.0;
    int4 it = make_int4(0, 0, 0, 0), h = it;
    if (warp < n_items) { it = __ldg(&items[warp]); h = __ldg(&ihdr[warp]); }
    for (int i = warp; i < n_items; i += n_warps) {
        int4 itn = make_int4(0, 0, 0, 0), hn = itn;                  // prefetch the next item of this warp
        if (i + n_warps < n_items) { itn = __ldg(&items[i + n_warps]); hn = __ldg(&ihdr[i + n_warps]); }
        const int len = it.y;
        const int2* __restrict__ cp = contacts + (((long long)it.w << 32) | (unsigned)it.z);
        const int r_id = h.x; const float r_mid = __int_as_float(h.y);
        const int wlo = all_w ? 0 : h.z; const unsigned wspan = all_w ? (unsigned)W : ((unsigned)h.w & 0x7fffffffu);
        if (h.w >= 0) {
            const int idle = wlo - 1;
            cp += lane;
            for (int off = 0; off < len; off += 32 * UN) {
                int2 ce[UN];
                const int left = len - off - lane;
                #pragma unroll
                for (int u = 0; u < UN; u++) ce[u] = (u * 32 < left) ? ld_stream(cp + off + u * 32) : make_int2(idle, 0);
                #pragma unroll
                for (int g = 0; g < UN / SUB; g++) {
                    bool anyw = false;
                    #pragma unroll
                    for (int j = 0; j < SUB; j++) anyw |= (unsigned)(ce[g * SUB + j].x - wlo) <= wspan;
                    if (!__any_sync(0xffffffffu, anyw)) continue;
                    int2 pc[SUB];
                    #pragma unroll
                    for (int j = 0; j < SUB; j++) {
                        pc[j] = make_int2(-1, 0);                    // (-1 outside the window: never a contig id)
                        if ((unsigned)(ce[g * SUB + j].x - wlo) <= wspan) pc[j] = __ldg(&cm[ce[g * SUB + j].x]);
                    }
                    unsigned bb[SUB]; int4 e[SUB]; unsigned fastm = 0u, slow = 0u;
                    #pragma unroll
                    for (int j = 0; j < SUB; j++) {
                        bb[j] = __float_as_uint(fabsf(__int_as_float(pc[j].y) - r_mid));
                        const unsigned t = bb[j] - smin;
                        const bool cis = pc[j].x == r_id;
                        const bool fast = cis && t < span;
                        e[j] = STAB ? stab[STAB ? ((fast ? t : 0u) >> (23 - TM)) : 0] : __ldg(&tab[(fast ? t : 0u) >> (23 - TM)]);
                        if (fast) fastm |= 1u << j;
                        else if (cis && ((bb[j] - 1u) < (smin - 1u) || (bb[j] - zlo) < zspan)) slow |= 1u << j;
                    }
                    float accf = 0.0f;
                    #pragma unroll
                    for (int j = 0; j < SUB; j++) {
                        const float uu = __uint_as_float(0x3f800000u | ((bb[j] & ((1u << (23 - TM)) - 1u)) << TM));
                        const float ob = ((fastm >> j) & 1u) ? __int_as_float(ce[g * SUB + j].y) : 0.0f;
                        accd = fma((double)ob, __hiloint2double(e[j].y, e[j].x), accd);
                        accf = fmaf(ob, uu * fmaf(uu, __int_as_float(e[j].w), __int_as_float(e[j].z)), accf);
                    }
                    accd += (double)accf;
                    if (slow) {                                      // in-band distances outside the table (rare)
                        #pragma unroll
                        for (int j = 0; j < SUB; j++)
                            if ((slow >> j) & 1u) accd += win_slow_entry(ce[g * SUB + j].x, __int_as_float(ce[g * SUB + j].y), r_mid, cm, p, lg);
                    }
                }
            }
        } else accd += win_circular_item(cp, len, lane, wlo, wspan, r_id, r_mid, __ldg(&geo[it.x].stot), cm, p, lg);
        it = itn; h = hn;
    }
    accd = block_sum(accd);
    if (threadIdx.x == 0) partials[blockIdx.x] = accd;
}

// (entries touching a duplicated data bin are left to the repeat path: sub_dup marks their sub-frags;
//  row_of gives the row of each 256-entry group to walk from)
__device__ __forceinline__ bool entry_excluded(const long long* __restrict__ rowptr, int W, const unsigned char* __restrict__ sub_dup,
                                               long long e, int col) {
    if (!sub_dup) return false;
    if (sub_dup[col]) return true;
    int lo = 0, hi = W;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (rowptr[mid + 1] > e) hi = mid; else lo = mid + 1; }
    return sub_dup[lo] != 0;
}
// sum of ob over all stored contacts (level constant)
__global__ void k_ob_total(const int2* __restrict__ contacts, long long E, const long long* __restrict__ rowptr, int W,
                           const unsigned char* __restrict__ sub_dup, double* __restrict__ partials) {
    double acc = 0.0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x)
        if (!entry_excluded(rowptr, W, sub_dup, e, contacts[e].x)) acc += (double)__int_as_float(contacts[e].y);
    acc = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// sum of lf(ob) over all stored contacts (level constant)
__global__ void k_lf_total(const int2* __restrict__ contacts, long long E, const long long* __restrict__ rowptr, int W,
                           const unsigned char* __restrict__ sub_dup, double* __restrict__ partials) {
    double acc = 0.0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x)
        if (!entry_excluded(rowptr, W, sub_dup, e, contacts[e].x)) acc += log_fact_term(__int_as_float(contacts[e].y));
    acc = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// ------------------------------------------------------------------------------------------------
// position order of bins (contig by contig) and the band-limited expected mass
// ------------------------------------------------------------------------------------------------
__global__ void k_contig_lengths(const int* __restrict__ slot, int ld, int n, int cap, int* __restrict__ cont_len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (slot[F_POS * ld + i] == 0) { const int c = slot[F_ID_C * ld + i]; if (c >= 0 && c < cap) cont_len[c] = slot[F_L_CONT * ld + i]; }
}
__global__ void k_order_fill(const int* __restrict__ slot, int ld, int n, int cap, const int* __restrict__ cont_off,
                             int* __restrict__ order) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = slot[F_ID_C * ld + i];
    if (c < 0 || c >= cap) return;
    const int k = cont_off[c] + slot[F_POS * ld + i];
    if (k >= 0 && k < n) order[k] = i;
}

// B(S) of a whole slot: sum over cis sub-frag pairs within d_max of  ex - g(true product).
// One warp per bin x (in position order); lanes take the following bins y of the same contig.  Cross-bin pairs
// -> partials[0][.], same-bin pairs a < b (the diagonal pixels, Q4) -> partials[1][.].  Runs when the cached
// band total is (re)synchronised; the step path updates the total from the committed candidate's band delta.
__global__ void __launch_bounds__(256)
k_band(const int* __restrict__ order, int count, const int* __restrict__ slot, int ld, LevelView lv, const Geo* __restrict__ geo,
       const __grid_constant__ Params p, double* __restrict__ partials, int partial_stride) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    double acc = 0.0, acc_diag = 0.0;
    for (int ix = warp; ix < count; ix += n_warps) {
        const int x = order[ix];
        if (x < 0 || !eligible(lv, x)) continue;
        const int4 sx = lv.sub_id[slot[F_ID_D * ld + x]];
        Geo gx[3];
        float xmax = -1e30f;
        #pragma unroll
        for (int a = 0; a < 3; a++) if (a < sx.w) { gx[a] = ld_geo(&geo[sx.x + a]); xmax = fmaxf(xmax, gx[a].mid); }
        const int cx = gx[0].id_c;
        if (lane == 0) {                             // same-bin pairs a < b (diagonal pixel, Q4)
            #pragma unroll
            for (int a = 0; a < 3; a++) for (int b = a + 1; b < 3; b++) if (b < sx.w) {
                const float s = fabsf(gx[b].mid - gx[a].mid);
                if (s > 0.0f && s < p.d_max) acc_diag += band_excess(gx[a], gx[b], s, p);
            }
        }
        for (int base = ix + 1; base < count; base += 32) {
            const int iy = base + lane;
            bool live = iy < count;
            if (live) {
                const int y = max(order[iy], 0);
                const float ystart = __int2float_rn(slot[F_START_BP * ld + y]) / 1000.0f;
                const int4 sy = lv.sub_id[slot[F_ID_D * ld + y]];
                const Geo gy0 = ld_geo(&geo[sy.x]);
                // beyond the band (or next contig): every remaining pair evaluates to the clamp value
                if (gy0.id_c != cx || (double)ystart - (double)xmax > (double)p.d_max * 1.00001 + 0.05) live = false;
                else if (eligible(lv, y)) {
                    #pragma unroll
                    for (int b = 0; b < 3; b++) if (b < sy.w) {
                        const Geo gy = (b == 0) ? gy0 : ld_geo(&geo[sy.x + b]);
                        #pragma unroll
                        for (int a = 0; a < 3; a++) if (a < sx.w) {
                            const float s = fabsf(gy.mid - gx[a].mid);
                            if (s > 0.0f && s < p.d_max) acc += band_excess(gx[a], gy, s, p);
                        }
                    }
                }
            }
            if (!__any_sync(0xffffffffu, live)) break;
        }
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
    acc_diag = block_sum(acc_diag);
    if (threadIdx.x == 0) partials[(size_t)1 * partial_stride + blockIdx.x] = acc_diag;
}

// same-bin pairs only (the diagonal pixels): thread per bin
__global__ void __launch_bounds__(256)
k_band_diag(const int* __restrict__ slot, int ld, int n, LevelView lv, const Geo* __restrict__ geo,
            const __grid_constant__ Params p, double* __restrict__ partials) {
    double acc = 0.0;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < n; x += gridDim.x * blockDim.x) {
        if (!eligible(lv, x)) continue;
        const int4 sx = lv.sub_id[slot[F_ID_D * ld + x]];
        Geo gx[3];
        #pragma unroll
        for (int a = 0; a < 3; a++) if (a < sx.w) gx[a] = ld_geo(&geo[sx.x + a]);
        #pragma unroll
        for (int a = 0; a < 3; a++) for (int b = a + 1; b < 3; b++) if (b < sx.w) {
            const float s = fabsf(gx[b].mid - gx[a].mid);
            if (s > 0.0f && s < p.d_max) acc += band_excess(gx[a], gx[b], s, p);
        }
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// Q(S): quirk-Q1 correction of trans pairs (bi < bj) whose lower bin is flipped and has non-uniform accu.
// One block per quirky bin of the list; bins_u == nullptr: all bins j > bi, else the members of bins_u.
__global__ void __launch_bounds__(256)
k_quirk(const int* __restrict__ quirky, int n_quirky, const int* __restrict__ slot, int ld, int n,
        size_t cand_slot_stride, LevelView lv, const int* __restrict__ bins_u, const int* __restrict__ d_count_u,
        const __grid_constant__ Params p, double* __restrict__ partials, int partial_stride) {
    const int k = blockIdx.y;
    const int* sl = slot + (size_t)k * cand_slot_stride;
    const int bi = quirky[blockIdx.x];
    double acc = 0.0;
    const bool act = sl[F_ORI * ld + bi] != 1;
    if (act) {
        const int cnt = bins_u ? *d_count_u : n;
        const int ci = sl[F_ID_C * ld + bi];
        const int4 si = lv.sub_id[bi];
        const int li = si.w - 1;
        const int ai[3] = { lv.sub_accu[bi * 3], lv.sub_accu[bi * 3 + 1], lv.sub_accu[bi * 3 + 2] };
        const int alast = ai[li];
        bool in_u = bins_u == nullptr;
        if (bins_u) { for (int j = threadIdx.x; j < cnt; j += blockDim.x) if (bins_u[j] == bi) in_u = true; in_u = __syncthreads_or(in_u); }
        if (in_u) {
            for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
                const int bj = bins_u ? bins_u[j] : j;
                if (bj <= bi || !eligible(lv, bj) || sl[F_ID_C * ld + bj] == ci) continue;
                const int lj = lv.sub_id[bj].w - 1;
                for (int b = 0; b <= lj; b++) {
                    const int aj = lv.sub_accu[bj * 3 + b];
                    for (int a = 0; a <= li; a++)
                        acc += (double)g_clamp(alast * aj, p.v_inter, p.nfpb) - (double)g_clamp(ai[a] * aj, p.v_inter, p.nfpb);
                }
            }
        }
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) partials[(size_t)k * partial_stride + blockIdx.x] = acc;
}

// ------------------------------------------------------------------------------------------------
// delta likelihood of candidate slots against the base slot (sub_compute_likelihood, range 1)
// ------------------------------------------------------------------------------------------------
// meta[0]=contig A, [1]=contig B, [2]=len A, [3]=len B (0 if same contig), [4]=m = |U|, [5]=max_id
// Start of a proposal's delta, one launch: the touched contigs (meta), U in the base slot's position order
// (fill_sub_index_fA / fB, kernels3.cu:3225-3249), and the per-proposal scratch cleared (chmask, piece lengths,
// changed ranges).  Every thread derives the contigs itself; thread 0 publishes them for the later kernels.
//   meta: [0] contig of fA, [1] contig of fB, [2] |A|, [3] |B| (0 if the same contig), [4] |U|, [5] max contig id
__global__ void k_delta_setup(const int* __restrict__ base, int ld, int fA, int fB, const int* __restrict__ d_max_id,
                              int max_id_host, int* __restrict__ meta, int n, int W, int* __restrict__ sub_index,
                              unsigned* __restrict__ chmask, int* __restrict__ piece_len, int* __restrict__ rng, unsigned* __restrict__ chmask2) {
    const int cA = base[F_ID_C * ld + fA], cB = base[F_ID_C * ld + fB];
    const int lA = base[F_L_CONT * ld + fA], lB = (cB != cA) ? base[F_L_CONT * ld + fB] : 0;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { meta[0] = cA; meta[1] = cB; meta[2] = lA; meta[3] = lB; meta[4] = lA + lB; meta[5] = (max_id_host >= 0) ? max_id_host : *d_max_id;
                  meta[6] = (fA == fB) ? 1 : 0; }        // [6] != 0: the general contact kernel scores this proposal (degenerate proposal / circular contig)
    if (i < GRAAL_N_CANDIDATES * 8) piece_len[i] = 0;
    if (i < 2 + 2 * N_BAND_ROWS) rng[i] = (i & 1) ? -1 : INT_MAX;
    for (int j = i; j < W; j += gridDim.x * blockDim.x) { chmask[j] = 0u; chmask2[j] = 0u; }
    if (i >= n) return;
    const int c = base[F_ID_C * ld + i];
    const int pos = base[F_POS * ld + i];
    if (c == cA) { if (pos >= 0 && pos < n) sub_index[pos] = i; }
    else if (c == cB) { const int k = lA + pos; if (k >= 0 && k < n) sub_index[k] = i; }
}
// profiler only: size of U of this proposal -> counters [0] entries of the rows of U, [1] rows (sub-frags), [2] bins, [3] proposals
__global__ void k_u_stats(const int* __restrict__ sub_index, const int* __restrict__ meta, LevelView lv, const int* __restrict__ base, int ld,
                          const long long* __restrict__ rowptr, unsigned long long* __restrict__ counters) {
    const int m = meta[4];
    unsigned long long e = 0, r = 0, b = 0;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < m; u += gridDim.x * blockDim.x) {
        const int4 sid = lv.sub_id[base[F_ID_D * ld + sub_index[u]]];
        e += (unsigned long long)(rowptr[sid.x + sid.w] - rowptr[sid.x]); r += sid.w; b += 1;
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) { e += __shfl_down_sync(0xffffffffu, e, o); r += __shfl_down_sync(0xffffffffu, r, o); b += __shfl_down_sync(0xffffffffu, b, o); }
    if ((threadIdx.x & 31) == 0 && b) { atomicAdd(&counters[0], e); atomicAdd(&counters[1], r); atomicAdd(&counters[2], b); }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&counters[3], 1ull);
}
// candidate geometry (members of U only) + contig lengths of the candidate's pieces of U
__device__ __forceinline__ int piece_slot(int id_c, const int* meta) {
    if (id_c == meta[0]) return 0;
    if (id_c == meta[1]) return 1;
    const int d = id_c - meta[5];
    return (d >= 1 && d <= 3) ? 1 + d : -1;
}
__global__ void k_cand_geometry(const int* __restrict__ cand0, size_t slot_stride, int ld, LevelView lv,
                                const int* __restrict__ sub_index, const int* __restrict__ meta,
                                Geo* __restrict__ geo0, size_t geo_stride, int* __restrict__ piece_len,
                                const Geo* __restrict__ geo_base, unsigned* __restrict__ chmask, unsigned skip_cands, int* __restrict__ meta_flag,
                                int4* __restrict__ rec_a0, int order_stride) {
    const int k = blockIdx.y;
    if ((skip_cands >> k) & 1u) return;
    const int m = meta[4];
    const int* sl = cand0 + (size_t)k * slot_stride;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < m; u += gridDim.x * blockDim.x) {
        // order records of this proposal start empty (n_sub = 0: skipped by the band pass): a degenerate proposal
        // (fA == fB) leaves order positions unwritten, which must not show the records of an earlier proposal
        rec_a0[(size_t)k * order_stride + u] = make_int4(0, 0, 0, -1);
        if (k < 3) rec_a0[(size_t)(GRAAL_N_CANDIDATES + k) * order_stride + u] = make_int4(0, 0, 0, -1);
        const int bin = sub_index[u];
        bin_geometry(sl, ld, bin, lv, geo0 + (size_t)k * geo_stride);
        if (sl[F_CIRC * ld + bin] == 1) meta_flag[0] = 1;
        if (eligible(lv, bin)) {   // bit k of chmask[sub]: the candidate's record differs from the base slot's (bitwise)
            const int4 sid = lv.sub_id[sl[F_ID_D * ld + bin]];
            for (int a = 0; a < sid.w; a++)
                if (!geo_eq(geo0[(size_t)k * geo_stride + sid.x + a], geo_base[sid.x + a])) atomicOr(&chmask[sid.x + a], 1u << k);
        }
        if (sl[F_POS * ld + bin] == 0) {
            const int ps = piece_slot(sl[F_ID_C * ld + bin], meta);
            if (ps >= 0) piece_len[k * 8 + ps] = sl[F_L_CONT * ld + bin];
        }
    }
}
// Position order of U in every candidate + 16-byte records per order position (what the band pass reads,
// coalesced, instead of chasing order -> slot -> sub_id -> geometry per neighbour):
//   A: x = first sub-frag | n_sub << 28 (n_sub = 0: duplicated bin, not scored here), y = start of the bin in kb
//      (float bits), z = bit a set if the record of sub-frag a differs from the base slot in this candidate,
//      w = contig id
//   B: x, y, z = mid-points of the sub-frags (float bits), w = accu index a | b << 8 | c << 16 | circular << 24
// and the first / last order index of a changed bin (rng[2k], rng[2k+1]).
__device__ __forceinline__ int4 band_record_b(const Geo* __restrict__ g, int sub0, int n_sub) {
    int mid[3] = {0, 0, 0}; unsigned info = 0u;
    for (int a = 0; a < n_sub; a++) {
        const Geo r = g[sub0 + a];
        mid[a] = __float_as_int(r.mid);
        info |= (unsigned)pk_true(r.pk) << (8 * a);
        if (a == 0) info |= (unsigned)pk_circ(r.pk) << 24;
    }
    return make_int4(mid[0], mid[1], mid[2], (int)info);
}
// the same records for the base slot in ITS order (sub_index); A.z = OR of the chmask words of the bin, and
//   C: x, y, z = chmask words of the sub-frags (bit k: differs from the base slot in candidate k)
__device__ __forceinline__ void base_order(const int* __restrict__ base, int ld, const int* __restrict__ sub_index, const int* __restrict__ meta,
                             LevelView lv, const unsigned* __restrict__ chmask, const Geo* __restrict__ geo_base,
                             int4* __restrict__ rec_a, int4* __restrict__ rec_b, int4* __restrict__ rec_c, int* __restrict__ rng, unsigned pair_mask,
                             int* __restrict__ meta_flag, const int4* __restrict__ row_hdr, int2* __restrict__ uwin) {
    const int m = meta[4];
    int lo = INT_MAX, hi = -1;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < m; u += gridDim.x * blockDim.x) {
        const int bin = sub_index[u];
        const int4 sid = lv.sub_id[base[F_ID_D * ld + bin]];
        const int n_sub = eligible(lv, bin) ? sid.w : 0;
        unsigned ms[3] = {0u, 0u, 0u};
        for (int a = 0; a < n_sub; a++) ms[a] = chmask[sid.x + a] & ~pair_mask;      // paired candidates get no old values: they start from their partner
        const unsigned chg = ms[0] | ms[1] | ms[2];
        const float start_kb = __int2float_rn(base[F_START_BP * ld + bin]) / 1000.0f;
        rec_a[u] = make_int4(sid.x | (n_sub << 28), __float_as_int(start_kb), (int)chg, base[F_ID_C * ld + bin]);
        rec_b[u] = band_record_b(geo_base, sid.x, n_sub);
        rec_c[u] = make_int4((int)ms[0], (int)ms[1], (int)ms[2], 0);
        if (base[F_CIRC * ld + bin] == 1) meta_flag[0] = 1;
        if (uwin) { const int4 h = row_hdr[sid.x]; uwin[sid.x] = make_int2(h.z, h.z + (int)((unsigned)h.w & 0x7fffffffu)); }   // union window starts as the base window
        if (chg) { lo = min(lo, u); hi = max(hi, u); }
    }
    if (hi >= 0) { atomicMin(&rng[0], lo); atomicMax(&rng[1], hi); }
}
// grid.y: rows 0..12 = candidates, row 13 = the base slot, rows 14..16 = the partners (2, 4, 6) of the paired
// candidates (3, 5, 7) restricted to the records in which the pair differs -> band rows 13..15.
//   normal candidate k : changed bits = chmask bit k (differs from the base slot)
//   paired candidate k : changed bits = differs from candidate k - 1 (also published in chmask2 bit k)
//   partner row of k   : candidate k - 1's order and records, the same changed bits
__global__ void k_cand_order(const int* __restrict__ cand0, size_t slot_stride, int ld,
                             const int* __restrict__ sub_index, const int* __restrict__ meta,
                             const int* __restrict__ piece_len, int order_stride, unsigned skip_cands,
                             LevelView lv, const unsigned* __restrict__ chmask, const Geo* __restrict__ geo0, size_t geo_stride,
                             int4* __restrict__ rec_a0, int4* __restrict__ rec_b0, int* __restrict__ rng,
                             int n_cand, const int* __restrict__ base, const Geo* __restrict__ geo_base,
                             int4* __restrict__ base_a, int4* __restrict__ base_b, int4* __restrict__ base_c, int* __restrict__ base_rng,
                             unsigned pair_mask, unsigned* __restrict__ chmask2, int* __restrict__ meta_flag, const int4* __restrict__ row_hdr, int2* __restrict__ uwin) {
    const int y = blockIdx.y;
    if (y == n_cand) { base_order(base, ld, sub_index, meta, lv, chmask, geo_base, base_a, base_b, base_c, base_rng, pair_mask, meta_flag, row_hdr, uwin); return; }
    int row = y, src = y, other = -1;                       // output row, candidate whose order / records are written, candidate compared with
    if (y > n_cand) {                                       // partner row v of paired candidate 3 + 2 v
        const int v = y - n_cand - 1, kp = 3 + 2 * v;
        if (!((pair_mask >> kp) & 1u)) return;
        row = GRAAL_N_CANDIDATES + v; src = kp - 1; other = kp;
    } else {
        if ((skip_cands >> y) & 1u) return;
        if ((pair_mask >> y) & 1u) other = y - 1;
    }
    const int m = meta[4];
    const int* sl = cand0 + (size_t)src * slot_stride;
    const Geo* gs = geo0 + (size_t)src * geo_stride;
    const Geo* go = (other >= 0) ? geo0 + (size_t)other * geo_stride : nullptr;
    int off[5]; int run = 0;
    #pragma unroll
    for (int s = 0; s < 5; s++) { off[s] = run; run += piece_len[src * 8 + s]; }
    int lo = INT_MAX, hi = -1;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < m; u += gridDim.x * blockDim.x) {
        const int bin = sub_index[u];
        const int ps = piece_slot(sl[F_ID_C * ld + bin], meta);
        if (ps < 0) continue;
        const int idx = off[ps] + sl[F_POS * ld + bin];
        if (idx < 0 || idx >= m) continue;
        const int4 sid = lv.sub_id[sl[F_ID_D * ld + bin]];
        const int n_sub = eligible(lv, bin) ? sid.w : 0;
        unsigned chg = 0u;
        for (int a = 0; a < n_sub; a++) {
            const int sub = sid.x + a;
            unsigned d;
            if (other < 0) d = (chmask[sub] >> src) & 1u;
            else {
                // a record not flagged in chmask was not rewritten for this proposal: it equals the base record
                const Geo rs = ((chmask[sub] >> src) & 1u) ? gs[sub] : geo_base[sub];
                const Geo ro = ((chmask[sub] >> other) & 1u) ? go[sub] : geo_base[sub];
                d = geo_eq(rs, ro) ? 0u : 1u;
                if (d && y < n_cand) atomicOr(&chmask2[sub], 1u << y);
            }
            chg |= d << a;
        }
        const float start_kb = __int2float_rn(sl[F_START_BP * ld + bin]) / 1000.0f;
        rec_a0[(size_t)row * order_stride + idx] = make_int4(sid.x | (n_sub << 28), __float_as_int(start_kb), (int)chg, sl[F_ID_C * ld + bin]);
        rec_b0[(size_t)row * order_stride + idx] = band_record_b(gs, sid.x, n_sub);
        if (chg) { lo = min(lo, idx); hi = max(hi, idx); }
    }
    if (hi >= 0) { atomicMin(&rng[2 * row], lo); atomicMax(&rng[2 * row + 1], hi); }
}

// ex - g of an in-band cis pair from its distance, accu table index and (circular contigs) contig length
__device__ __noinline__ double band_excess_general(float s, int idx, int circ, float stot, const Params& p) {
    Geo a; a.mid = 0.0f; a.id_c = 0; a.stot = stot; a.pk = (unsigned)(idx / p.nd) | ((unsigned)circ << 28);
    Geo b = a; b.pk = (unsigned)(idx % p.nd);
    return band_excess(a, b, s, p);
}

// Band mass of the delta from the ordered records: warp per bin x, lanes over the following bins.
//   BASE = false: candidate k = blockIdx.y, NEW values of pairs with a changed record (bit k)
//   BASE = true : base slot, OLD values once, credited to every candidate whose bit is set
// Only windows that can contain a changed pair are scanned: x beyond the last changed bin is skipped, an
// unchanged x starts its scan at the first changed bin and stops at the last one.  Per neighbour bin a lane
// reads two (three) coalesced records; the nine sub-frag pairs are evaluated branch-free on the tabulated
// law (one table gather each, nine independent chains); pairs outside the table, circular contigs and math
// modes 0 / 1 take the general path.
template <bool BASE, int PARTS, bool UNI>
__global__ void __launch_bounds__(256)
k_band_delta(const int4* __restrict__ rec_a0, const int4* __restrict__ rec_b0, const int4* __restrict__ rec_c, int order_stride,
             const int* __restrict__ d_count, const int* __restrict__ rng0,
             const Geo* __restrict__ geo0, size_t cand_geo_stride, unsigned skip_cands,
             const __grid_constant__ Params p, double* __restrict__ partials, int partial_stride) {
    const int k = blockIdx.y;
    const int count = *d_count;
    const int4* rec_a = BASE ? rec_a0 : rec_a0 + (size_t)k * order_stride;
    const int4* rec_b = BASE ? rec_b0 : rec_b0 + (size_t)k * order_stride;
    // band rows 13..15 hold the order / records of candidates 2, 4, 6 (partners of the paired candidates)
    const Geo* gE = BASE ? geo0 : geo0 + (size_t)(k < GRAAL_N_CANDIDATES ? k : 2 + 2 * (k - GRAAL_N_CANDIDATES)) * cand_geo_stride;
    const int* rng = BASE ? rng0 : rng0 + 2 * k;
    const int lo = rng[0], hi = rng[1];
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    // UNI: one accu value in the level (p.nd == 1), the pair tables are scalars
    const double normd0 = __ldg(&p.t_normd[0]), neg_g00 = -(double)__ldg(&p.t_g[0]);
    const int4* __restrict__ t_f = p.t_f;
    const double v_clamp = p.v_clamp;
    const float d_max = p.d_max;
    double acc = 0.0;
    double accs[GRAAL_N_CANDIDATES];
    if (BASE) {
        #pragma unroll
        for (int c = 0; c < GRAAL_N_CANDIDATES; c++) accs[c] = 0.0;
    }
    const bool idle = (!BASE && ((skip_cands >> k) & 1u)) || hi < 0;
    // PARTS warps share one x: warp part takes the 32-bin chunks part, part + PARTS, ... of the window
    if (!idle) for (int wx = warp; wx / PARTS <= hi && wx / PARTS < count; wx += n_warps) {
        const int ix = wx / PARTS, part = wx % PARTS;
        const int4 rx = rec_a[ix];
        const int nx = (rx.x >> 28) & 7;
        if (nx == 0) continue;                                   // duplicated bin: repeat path
        const unsigned xm = (unsigned)rx.z;
        const bool xchg = xm != 0u;
        int y0 = ix + 1, y1 = count;
        if (!xchg) { y0 = max(y0, lo); y1 = min(y1, hi + 1); }
        if (y0 >= y1) continue;
        const int4 xb = rec_b[ix];
        const float xmid[3] = {__int_as_float(xb.x), __int_as_float(xb.y), __int_as_float(xb.z)};
        unsigned mx[3];
        if (BASE) { const int4 xc = rec_c[ix]; mx[0] = (unsigned)xc.x; mx[1] = (unsigned)xc.y; mx[2] = (unsigned)xc.z; }
        else { mx[0] = xm & 1u; mx[1] = (xm >> 1) & 1u; mx[2] = (xm >> 2) & 1u; }
        float xmax = xmid[0];
        if (nx > 1) xmax = fmaxf(xmax, xmid[1]);
        if (nx > 2) xmax = fmaxf(xmax, xmid[2]);
        const int cx = rx.w;
        const int circ = (xb.w >> 24) & 1;
        const float stot = circ ? gE[rx.x & 0x0fffffff].stot : 0.0f;
        const bool fast_ok = p.mode == 2 && !circ;
        double tot[3] = {0.0, 0.0, 0.0};                         // BASE: old values of the pairs of sub a of x
        for (int basei = y0 + 32 * part; basei < y1; basei += 32 * PARTS) {
            const int iy = basei + lane;
            bool live = iy < y1;
            if (live) {
                const int4 ry = rec_a[iy];
                // beyond the band (or next contig): every remaining pair evaluates to the clamp value
                if (ry.w != cx || (double)__int_as_float(ry.y) - (double)xmax > (double)p.d_max * 1.00001 + 0.05) live = false;
                else {
                    const int ny = (ry.x >> 28) & 7;
                    const unsigned ym = (unsigned)ry.z;
                    if (ny > 0 && (xchg || ym != 0u)) {
                        const int4 yb = rec_b[iy];
                        const float ymid[3] = {__int_as_float(yb.x), __int_as_float(yb.y), __int_as_float(yb.z)};
                        unsigned my[3];
                        if (BASE) { const int4 yc = rec_c[iy]; my[0] = (unsigned)yc.x; my[1] = (unsigned)yc.y; my[2] = (unsigned)yc.z; }
                        else { my[0] = ym & 1u; my[1] = (ym >> 1) & 1u; my[2] = (ym >> 2) & 1u; }
                        unsigned slow = 0u;                        // pairs left to the general path (rare): after the fast ones
                        #pragma unroll
                        for (int b = 0; b < 3; b++) {
                            #pragma unroll
                            for (int a = 0; a < 3; a++) {
                                const unsigned mm = mx[a] | my[b];
                                const float s = fabsf(ymid[b] - xmid[a]);
                                const bool on = a < nx && b < ny && mm != 0u && s > 0.0f && s < d_max;
                                const bool fast = on && fast_ok && law_in_table(s);
                                if (on && !fast) slow |= 1u << (3 * b + a);
                                const int idx = UNI ? 0 : (int)((xb.w >> (8 * a)) & 255) * p.nd + (int)((yb.w >> (8 * b)) & 255);
                                const double f = law_interp(fast ? s : 1.0f, t_f);
                                const double vf = UNI ? fma(f, normd0, neg_g00) : fma(f, __ldg(&p.t_normd[idx]), -(double)__ldg(&p.t_g[idx]));
                                const double v = (fast && f > v_clamp) ? vf : 0.0;
                                if (BASE) {
                                    tot[a] += v;                              // credited below to every candidate of mx[a]
                                    const unsigned extra = fast ? (my[b] & ~mx[a]) : 0u;
                                    if (extra) {
                                        #pragma unroll
                                        for (int c = 0; c < GRAAL_N_CANDIDATES; c++) if ((extra >> c) & 1u) accs[c] += v;
                                    }
                                } else acc += v;
                            }
                        }
                        if (slow) {
                            #pragma unroll
                            for (int b = 0; b < 3; b++) {
                                #pragma unroll
                                for (int a = 0; a < 3; a++) {
                                    if (!((slow >> (3 * b + a)) & 1u)) continue;
                                    const int idx = UNI ? 0 : (int)((xb.w >> (8 * a)) & 255) * p.nd + (int)((yb.w >> (8 * b)) & 255);
                                    const double v = band_excess_general(fabsf(ymid[b] - xmid[a]), idx, circ, stot, p);
                                    if (BASE) {
                                        tot[a] += v;
                                        const unsigned extra = my[b] & ~mx[a];
                                        #pragma unroll
                                        for (int c = 0; c < GRAAL_N_CANDIDATES; c++) if ((extra >> c) & 1u) accs[c] += v;
                                    } else acc += v;
                                }
                            }
                        }
                    }
                }
            }
            if (!__any_sync(0xffffffffu, live)) break;
        }
        if (BASE) {
            #pragma unroll
            for (int a = 0; a < 3; a++) {
                if (a >= nx || mx[a] == 0u) continue;
                #pragma unroll
                for (int c = 0; c < GRAAL_N_CANDIDATES; c++) if ((mx[a] >> c) & 1u) accs[c] += tot[a];
            }
        }
    }
    if (BASE) {
        #pragma unroll
        for (int c = 0; c < GRAAL_N_CANDIDATES; c++) {
            const double v = block_sum(accs[c]);
            if (threadIdx.x == 0) partials[(size_t)c * partial_stride + blockIdx.x] = v;
        }
    } else {
        acc = block_sum(acc);
        if (threadIdx.x == 0) partials[(size_t)k * partial_stride + blockIdx.x] = acc;
    }
}

// The same band pass for uniform-accu levels with a monotone tabulated law (Params::fu_ok), about half the instructions:
//  * the range test 0 < s < min(d_max, clamp threshold, table end) is ONE unsigned compare on the bit pattern of s, and
//    absent sub-frags carry a NaN mid-point, which fails it;
//  * the table already holds f(s) * norm - g: value = a0 + u * (a1 + u * a2), a0 accumulated in float64, the small
//    u-part in float32 (flushed per x);
//  * "pair has a changed record" is folded into the operands: against an unchanged sub-frag of x a lane uses the
//    mid-points of its CHANGED sub-frags only (the others NaN);
//  * BASE: the 13 per-candidate accumulators of a thread live in shared memory and are credited by looping over the set
//    bits of the candidate masks (a bin's sub-frags almost always share one mask: the nine pairs are summed first).
// Circular contigs and in-band distances outside the table take the general evaluation pair by pair.
struct FastBand { const int4* tab; unsigned smin_bits, span, zlo, zspan, dmax_bits; };
template <bool BASE, int PARTS>
__global__ void __launch_bounds__(256, 3)
k_band_delta_fast(const int4* __restrict__ rec_a0, const int4* __restrict__ rec_b0, const int4* __restrict__ rec_c, int order_stride,
                  const int* __restrict__ d_count, const int* __restrict__ rng0,
                  const Geo* __restrict__ geo0, size_t cand_geo_stride, unsigned skip_cands, const FastBand fb,
                  const __grid_constant__ Params p, double* __restrict__ partials, int partial_stride, int max_parts) {
    __shared__ double sacc[BASE ? GRAAL_N_CANDIDATES : 1][BASE ? 256 : 1];
    const int k = blockIdx.y;
    const int count = *d_count;
    const int4* rec_a = BASE ? rec_a0 : rec_a0 + (size_t)k * order_stride;
    const int4* rec_b = BASE ? rec_b0 : rec_b0 + (size_t)k * order_stride;
    const Geo* gE = BASE ? geo0 : geo0 + (size_t)(k < GRAAL_N_CANDIDATES ? k : 2 + 2 * (k - GRAAL_N_CANDIDATES)) * cand_geo_stride;
    const int* rng = BASE ? rng0 : rng0 + 2 * k;
    const int lo = rng[0], hi = rng[1];
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int4* __restrict__ tab = fb.tab;
    const float d_max = p.d_max;
    const float nanf_ = __int_as_float(0x7fc00000);
    double acc = 0.0;
    if (BASE) {
        #pragma unroll
        for (int c = 0; c < GRAAL_N_CANDIDATES; c++) sacc[c][threadIdx.x] = 0.0;
    }
    const bool idle = (!BASE && ((skip_cands >> k) & 1u)) || hi < 0;
    // a short U leaves most warps without a bin and the others with a chain of dependent trips along the band of theirs:
    // the band of every x is cut over `parts` warps (a power of two >= PARTS) until every warp has work
    int parts = PARTS;
    if (max_parts > PARTS) { const int nx_ = min(hi + 1, count); while (parts < max_parts && nx_ * parts * 2 <= n_warps) parts <<= 1; }
    if (!idle) for (int wx = warp; wx / parts <= hi && wx / parts < count; wx += n_warps) {
        const int ix = wx / parts, part = wx - ix * parts;
        const int4 rx = rec_a[ix];
        const int nx = (rx.x >> 28) & 7;
        if (nx == 0) continue;                                   // duplicated bin: repeat path
        const unsigned xm = (unsigned)rx.z;
        const bool xchg = xm != 0u;
        int y0 = ix + 1, y1 = count;
        if (!xchg) { y0 = max(y0, lo); y1 = min(y1, hi + 1); }
        if (y0 >= y1) continue;
        const int4 xb = rec_b[ix];
        unsigned mx[3];
        if (BASE) { const int4 xc = rec_c[ix]; mx[0] = (unsigned)xc.x; mx[1] = (unsigned)xc.y; mx[2] = (unsigned)xc.z; }
        else { mx[0] = xm & 1u; mx[1] = (xm >> 1) & 1u; mx[2] = (xm >> 2) & 1u; }
        float xmid[3] = {__int_as_float(xb.x), nx > 1 ? __int_as_float(xb.y) : nanf_, nx > 2 ? __int_as_float(xb.z) : nanf_};
        float xmax = xmid[0];
        if (nx > 1) xmax = fmaxf(xmax, xmid[1]);
        if (nx > 2) xmax = fmaxf(xmax, xmid[2]);
        const int cx = rx.w;
        const int circ = (xb.w >> 24) & 1;
        const float stot = circ ? gE[rx.x & 0x0fffffff].stot : 0.0f;
        // circular contig: nothing is fast, every in-band pair goes to the general evaluation
        const unsigned span = circ ? 0u : fb.span, zlo = circ ? 1u : fb.zlo, zspan = circ ? fb.dmax_bits - 1u : fb.zspan, smin = fb.smin_bits;
        const bool x_uniform = (nx < 2 || mx[1] == mx[0]) && (nx < 3 || mx[2] == mx[0]);
        double tot = 0.0; float totf = 0.0f, accf = 0.0f;        // BASE: old values credited to the candidates of the whole bin x
        // the records of the NEXT chunk of 32 bins are loaded before the current chunk is evaluated (the three dependent
        // stages record A -> record B -> table would otherwise be paid per chunk with a handful of warps per scheduler)
        const int4 zero4 = make_int4(0, 0, 0, 0);
        int iy0 = y0 + 32 * part + lane;
        int4 ry_n = iy0 < y1 ? rec_a[iy0] : zero4, yb_n = iy0 < y1 ? rec_b[iy0] : zero4, yc_n = (BASE && iy0 < y1) ? rec_c[iy0] : zero4;
        for (int basei = y0 + 32 * part; basei < y1; basei += 32 * parts) {
            const int iy = basei + lane;
            bool live = iy < y1;
            const int4 ry = ry_n, yb = yb_n, yc4 = yc_n;
            {
                const int in = iy + 32 * parts;
                ry_n = in < y1 ? rec_a[in] : zero4; yb_n = in < y1 ? rec_b[in] : zero4;
                if (BASE) yc_n = in < y1 ? rec_c[in] : zero4;
            }
            if (live) {
                // beyond the band (or next contig): every remaining pair evaluates to the clamp value
                if (ry.w != cx || (double)__int_as_float(ry.y) - (double)xmax > (double)d_max * 1.00001 + 0.05) live = false;
                else {
                    const int ny = (ry.x >> 28) & 7;
                    const unsigned ym = (unsigned)ry.z;
                    if (ny > 0 && (xchg || ym != 0u)) {
                        unsigned my[3];
                        if (BASE) { my[0] = (unsigned)yc4.x; my[1] = (unsigned)yc4.y; my[2] = (unsigned)yc4.z; }
                        else { my[0] = ym & 1u; my[1] = (ym >> 1) & 1u; my[2] = (ym >> 2) & 1u; }
                        const float ya[3] = {__int_as_float(yb.x), ny > 1 ? __int_as_float(yb.y) : nanf_, ny > 2 ? __int_as_float(yb.z) : nanf_};
                        const float yc[3] = {my[0] ? ya[0] : nanf_, my[1] ? ya[1] : nanf_, my[2] ? ya[2] : nanf_};   // changed sub-frags only
                        const bool uniform = BASE && x_uniform && (ny < 2 || my[1] == my[0]) && (ny < 3 || my[2] == my[0]);
                        unsigned slow = 0u;
                        double pd = 0.0; float pf = 0.0f;                // sum of the nine pairs
                        #pragma unroll
                        for (int b = 0; b < 3; b++) {
                            #pragma unroll
                            for (int a = 0; a < 3; a++) {
                                const float yv = mx[a] ? ya[b] : yc[b];
                                const unsigned sb = __float_as_uint(fabsf(yv - xmid[a]));
                                const unsigned t = sb - smin;
                                const bool fast = t < span;
                                if ((sb - 1u) < (smin - 1u) || (sb - zlo) < zspan) slow |= 1u << (3 * b + a);
                                const int4 e = __ldg(&tab[(fast ? t : 0u) >> (23 - LAW_M)]);
                                const float uu = __uint_as_float(0x3f800000u | ((sb & ((1u << (23 - LAW_M)) - 1u)) << LAW_M));
                                const double v0 = fast ? __hiloint2double(e.y, e.x) : 0.0;
                                const float v1 = fast ? uu * fmaf(uu, __int_as_float(e.w), __int_as_float(e.z)) : 0.0f;
                                if (!BASE || uniform) { pd += v0; pf += v1; }
                                else {                                    // sub-frags of a bin with different candidate masks (rare)
                                    const double v = v0 + (double)v1;
                                    unsigned cm = mx[a] | my[b];
                                    while (cm) { const int c = __ffs(cm) - 1; cm &= cm - 1u; sacc[BASE ? c : 0][BASE ? threadIdx.x : 0] += v; }
                                }
                            }
                        }
                        if (slow) {
                            #pragma unroll
                            for (int b = 0; b < 3; b++) {
                                #pragma unroll
                                for (int a = 0; a < 3; a++) {
                                    if (!((slow >> (3 * b + a)) & 1u)) continue;
                                    const float yv = mx[a] ? ya[b] : yc[b];
                                    const float sv = fabsf(yv - xmid[a]);
                                    if (!(sv > 0.0f && sv < d_max)) continue;
                                    const double v = band_excess_general(sv, 0, circ, stot, p);
                                    if (!BASE || uniform) pd += v;
                                    else { unsigned cm = mx[a] | my[b]; while (cm) { const int c = __ffs(cm) - 1; cm &= cm - 1u; sacc[BASE ? c : 0][BASE ? threadIdx.x : 0] += v; } }
                                }
                            }
                        }
                        if (!BASE) { acc += pd; accf += pf; }
                        else if (uniform) {
                            tot += pd; totf += pf;                          // candidates of x: credited once per x below
                            unsigned extra = my[0] & ~mx[0];                // candidates in which only y changed
                            const double v = pd + (double)pf;
                            while (extra) { const int c = __ffs(extra) - 1; extra &= extra - 1u; sacc[BASE ? c : 0][BASE ? threadIdx.x : 0] += v; }
                        }
                    }
                }
            }
            if (!__any_sync(0xffffffffu, live)) break;
        }
        if (BASE) {
            const double v = tot + (double)totf;
            unsigned cm = mx[0];                                         // (x_uniform: the mask of the whole bin; else tot == 0)
            while (cm) { const int c = __ffs(cm) - 1; cm &= cm - 1u; sacc[BASE ? c : 0][BASE ? threadIdx.x : 0] += v; }
        } else acc += (double)accf;
    }
    if (BASE) {
        __syncthreads();
        #pragma unroll 1
        for (int c = 0; c < GRAAL_N_CANDIDATES; c++) {
            const double v = block_sum(sacc[BASE ? c : 0][BASE ? threadIdx.x : 0]);
            if (threadIdx.x == 0) partials[(size_t)c * partial_stride + blockIdx.x] = v;
        }
    } else {
        acc = block_sum(acc);
        if (threadIdx.x == 0) partials[(size_t)k * partial_stride + blockIdx.x] = acc;
    }
}

// Contact part of the delta, all candidates in ONE pass over the rows of U: one warp per sub-frag row of a
// member bin, four 32-entry slices of the row in flight per trip (the loads of a slice do not depend on
// the previous one).  For every stored contact whose partner is in U, in another bin, with a changed
// record on either side (chmask), the OLD term is evaluated once and the NEW term once per flagged
// candidate:  accs[k] += ob * (ln ex_k - ln ex_0).
// Uniform-accu levels with a monotone tabulated law: the contact term RELATIVE to the far value, ob * (ln ex - log g):
// exactly 0 for a far pair, one table gather (a0 holds ln norm - log g) for an in-band one; circular contigs and in-band
// distances outside the table take the general evaluation.
__device__ __forceinline__ double contact_rel_term(const Geo& a, const Geo& b, float ob, const FastLaw& fl, const Params& p, double lg) {
    const unsigned sb = __float_as_uint(fabsf(b.mid - a.mid));
    const unsigned t = sb - fl.smin_bits;
    const bool cis = a.id_c == b.id_c;
    const bool circ = pk_circ(a.pk) != 0;
    if (cis && t < fl.span && !circ) {
        const int4 e = __ldg(&fl.tab[t >> (23 - LAW_M)]);
        const float uu = __uint_as_float(0x3f800000u | ((sb & ((1u << (23 - LAW_M)) - 1u)) << LAW_M));
        return (double)ob * (__hiloint2double(e.y, e.x) + (double)(uu * fmaf(uu, __int_as_float(e.w), __int_as_float(e.z))));
    }
    if (cis && (circ || (sb - 1u) < (fl.smin_bits - 1u) || (sb - fl.zlo) < fl.zspan))
        return contact_log_term_general(a.mid, a.id_c, a.stot, a.pk, b.mid, b.id_c, b.pk, ob, p) - (double)ob * lg;
    return 0.0;
}

#define DC_UNROLL 2
#define DC_MAX_SPLIT 8
// MINB > 2: the 13 accumulators of a thread live in shared memory (26 KB per CTA) and the register budget is 64K / (256 MINB)
template <bool UNI, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_delta_contacts_rows(const long long* __restrict__ rowptr, const int2* __restrict__ contacts, LevelView lv,
                      const int* __restrict__ sub_index, const int* __restrict__ meta,
                      const Geo* __restrict__ geo_base, const Geo* __restrict__ geo_cand0, size_t geo_stride,
                      const unsigned* __restrict__ chmask, const unsigned* __restrict__ chmask2, unsigned pair_mask,
                      const __grid_constant__ Params p, double* __restrict__ partials, int partial_stride, const int2* __restrict__ uwin,
                      const FastLaw fl, double lg, int max_split) {
    const int m = meta[4], cA = meta[0], cB = meta[1];
    if (meta[6]) uwin = nullptr;                               // degenerate proposal (fA == fB): the candidate orders are not trusted
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    constexpr bool SACC = MINB > 2;
    __shared__ double sacc[SACC ? GRAAL_N_CANDIDATES : 1][SACC ? 256 : 1];
    double accs[SACC ? 1 : GRAAL_N_CANDIDATES];
    #pragma unroll
    for (int c = 0; c < GRAAL_N_CANDIDATES; c++) { if (SACC) sacc[SACC ? c : 0][SACC ? threadIdx.x : 0] = 0.0; else accs[SACC ? 0 : c] = 0.0; }
    // a short U leaves most warps without a row and the others with a chain of dependent trips over theirs: the rows are
    // cut into S slices (a power of two, whole trips each) until every warp has work or a slice is one trip
    int S = 1;
    while (S < max_split && 3 * m * S * 2 <= n_warps) S <<= 1;
    for (int w = warp; w < 3 * m * S; w += n_warps) {
        const int row = w / S, slice = w - row * S;
        const int u = row / 3, a = row - 3 * u;
        const int bin = sub_index[u];
        if (!eligible(lv, bin)) continue;                     // duplicated bin: repeat path
        const int4 sid = lv.sub_id[bin];
        if (a >= sid.w) continue;
        const int sub0 = sid.x, rowsub = sid.x + a;
        long long e0 = __ldg(&rowptr[rowsub]), e1 = __ldg(&rowptr[rowsub + 1]);
        if (S > 1) {
            const long long per = (((e1 - e0 + S - 1) / S) + 32 * DC_UNROLL - 1) / (32 * DC_UNROLL) * (32 * DC_UNROLL);
            e0 += slice * per;
            e1 = min(e1, e0 + per);
            if (e0 >= e1) continue;
        }
        const Geo r0 = ld_geo(&geo_base[rowsub]);
        const unsigned mra = __ldg(&chmask[rowsub]);
        const unsigned mra2 = pair_mask ? (__ldg(&chmask2[rowsub]) & pair_mask) : 0u;   // row differs between a paired candidate and its partner
        // union window of the row's bin (uniform levels): columns outside it are far in the base and in every candidate
        int wlo = 0; unsigned wspan = 0xffffffffu;
        if (uwin) { const int2 uw = __ldg(&uwin[sub0]); wlo = uw.x; wspan = (unsigned)(uw.y - uw.x); }
        for (long long eb = e0 + lane; eb < e1; eb += 32 * DC_UNROLL) {
            int2 ce[DC_UNROLL]; Geo g0[DC_UNROLL]; unsigned mc[DC_UNROLL]; bool ok[DC_UNROLL];
            #pragma unroll
            for (int j = 0; j < DC_UNROLL; j++) {
                ok[j] = eb + 32 * j < e1;
                ce[j] = ok[j] ? __ldg(&contacts[eb + 32 * j]) : make_int2(0, 0);   // rows of U recur in the next proposals: keep them in L2
                ok[j] = ok[j] && (unsigned)(ce[j].x - wlo) <= wspan;
            }
            #pragma unroll
            for (int j = 0; j < DC_UNROLL; j++) if (ok[j]) { g0[j] = ld_geo(&geo_base[ce[j].x]); mc[j] = __ldg(&chmask[ce[j].x]); }
            #pragma unroll
            for (int j = 0; j < DC_UNROLL; j++) {
                if (!ok[j]) continue;
                const Geo g0c = g0[j];
                if (g0c.id_c != cA && g0c.id_c != cB) continue;              // partner outside U
                if (g0c.pk & PK_EXCLUDED) continue;                          // partner is a duplicated bin: repeat path
                if (ce[j].x - pk_local(g0c.pk) == sub0) continue;            // same bin: diagonal pixel, not re-scored
                // paired candidates (bits of pair_mask) are scored against their partner: accs[k] = sum of
                // ob * (ln ex_k - ln ex_{k-1}) over the contacts with a record that differs between the two
                const unsigned mm = (mra | mc[j]) & ~pair_mask;
                const unsigned mm2 = pair_mask ? (mra2 | (__ldg(&chmask2[ce[j].x]) & pair_mask)) : 0u;
                if (!(mm | mm2)) continue;                                   // bitwise unchanged in every candidate
                const float ob = __int_as_float(ce[j].y);
                // {mid-point, contig id} of the partner in every flagged candidate (the rest of its record -- accu
                // indices, flags -- does not depend on the candidate): all gathers in flight before the first use
                int2 pr[GRAAL_N_CANDIDATES];
                #pragma unroll
                for (int k = 0; k < GRAAL_N_CANDIDATES; k++)
                    pr[k] = ((mc[j] >> k) & mm >> k & 1u) ? __ldg(reinterpret_cast<const int2*>(&geo_cand0[(size_t)k * geo_stride + ce[j].x]))
                                                          : make_int2(__float_as_int(g0c.mid), g0c.id_c);
                const double told = UNI ? contact_rel_term(r0, g0c, ob, fl, p, lg) : contact_log_term(r0, g0c, ob, p);
                #pragma unroll
                for (int k = 0; k < GRAAL_N_CANDIDATES; k++) {
                    if (!((mm >> k) & 1u)) continue;
                    const Geo rk = ((mra >> k) & 1u) ? ld_geo(&geo_cand0[(size_t)k * geo_stride + rowsub]) : r0;
                    Geo gc = g0c; gc.mid = __int_as_float(pr[k].x); gc.id_c = pr[k].y;
                    const double dv = (UNI ? contact_rel_term(rk, gc, ob, fl, p, lg) : contact_log_term(rk, gc, ob, p)) - told;
                    if (SACC) sacc[SACC ? k : 0][SACC ? threadIdx.x : 0] += dv; else accs[SACC ? 0 : k] += dv;
                }
                if (mm2) {                                                   // rare: the few records in which a pair differs
                    #pragma unroll
                    for (int k = 3; k <= 7; k += 2) {
                        if (!((mm2 >> k) & 1u)) continue;
                        double t[2];
                        #pragma unroll
                        for (int w2 = 0; w2 < 2; w2++) {                     // candidate k, then its partner k - 1
                            const int kk = k - w2;
                            const Geo rk = ((mra >> kk) & 1u) ? ld_geo(&geo_cand0[(size_t)kk * geo_stride + rowsub]) : r0;
                            Geo gc = g0c;
                            if ((mc[j] >> kk) & 1u) { const Geo q = ld_geo(&geo_cand0[(size_t)kk * geo_stride + ce[j].x]); gc.mid = q.mid; gc.id_c = q.id_c; }
                            t[w2] = UNI ? contact_rel_term(rk, gc, ob, fl, p, lg) : contact_log_term(rk, gc, ob, p);
                        }
                        if (SACC) sacc[SACC ? k : 0][SACC ? threadIdx.x : 0] += t[0] - t[1]; else accs[SACC ? 0 : k] += t[0] - t[1];
                    }
                }
            }
        }
    }
    #pragma unroll
    for (int c = 0; c < GRAAL_N_CANDIDATES; c++) {
        const double v = block_sum(SACC ? sacc[SACC ? c : 0][SACC ? threadIdx.x : 0] : accs[SACC ? 0 : c]);
        if (threadIdx.x == 0) partials[(size_t)c * partial_stride + blockIdx.x] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// Union windows of the delta contact pass (uniform-accu levels).
//
// On a uniform level a contact between two far sub-frags (trans, or cis beyond the band) evaluates to the same ob * log g
// in every structure, so a stored contact of U only matters for candidate k if the pair is in band in the base structure or
// in candidate k.  Per bin of U the UNION over the base and the 13 candidates of the sub-frag index hull of the bins within
// d_max (the base window from the windowed full pass, widened by every candidate's window, computed from the candidate's
// position-ordered records) bounds the columns that can matter: every other entry of the row costs its stream load and
// two integer ops in k_delta_contacts_rows -- no partner gather, no evaluation -- and contributes exactly 0 either way.
// ------------------------------------------------------------------------------------------------
// hull of the sub-frag index ranges of every aligned block of 32 order positions of every candidate (grid.y)
__global__ void k_cand_hull_blocks(const int4* __restrict__ rec_a0, int order_stride, const int* __restrict__ meta, int2* __restrict__ blk0, int blk_stride, unsigned skip_cands) {
    const int k = blockIdx.y;
    if ((skip_cands >> k) & 1u) return;
    const int m = meta[4];
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < ((m + 31) & ~31); g += gridDim.x * blockDim.x) {
        int lo = INT_MAX, hi = -1;
        if (g < m) { const int x = rec_a0[(size_t)k * order_stride + g].x; lo = x & 0x0fffffff; hi = lo + ((x >> 28) & 7) - 1; }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
        if ((threadIdx.x & 31) == 0) blk0[(size_t)k * blk_stride + (g >> 5)] = make_int2(lo, hi);
    }
}
// window of every bin of U in every candidate (grid.y), merged into the union window of the bin (uwin[first sub-frag]).
// reach_kb = d_max (+ 0.01 %) + the longest bin of the level + 0.1 kb: a bin further than that (start to start) holds no
// sub-frag within d_max.
__global__ void __launch_bounds__(256)
k_cand_windows(const int4* __restrict__ rec_a0, int order_stride, const int* __restrict__ meta, const int* __restrict__ piece_len,
               const int2* __restrict__ blk0, int blk_stride, unsigned skip_cands, float reach_kb, int2* __restrict__ uwin) {
    const int k = blockIdx.y;
    if ((skip_cands >> k) & 1u) return;
    const int m = meta[4];
    const int4* __restrict__ ra = rec_a0 + (size_t)k * order_stride;
    const int2* __restrict__ blk = blk0 + (size_t)k * blk_stride;
    int off[6]; int run = 0;
    #pragma unroll
    for (int q = 0; q < 5; q++) { off[q] = run; run += piece_len[k * 8 + q]; }
    off[5] = run;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < m; g += gridDim.x * blockDim.x) {
        const int4 r = ra[g];
        const int ps = piece_slot(r.w, meta);
        int cs = 0, ce = m;                                   // the piece (contig of the candidate) that holds position g
        if (ps >= 0) { cs = off[ps]; ce = min(off[ps + 1], m); }
        if (!(cs <= g && g < ce)) { cs = 0; ce = m; }
        const float st = __int_as_float(r.y);
        int a = cs, b = g;                                    // first position whose bin starts after st - reach
        while (a < b) { const int mm = (a + b) >> 1; if (__int_as_float(ra[mm].y) > st - reach_kb) b = mm; else a = mm + 1; }
        const int lo = a;
        a = g; b = ce - 1;                                    // last position whose bin starts before st + reach
        while (a < b) { const int mm = (a + b + 1) >> 1; if (__int_as_float(ra[mm].y) < st + reach_kb) a = mm; else b = mm - 1; }
        const int hi = a;
        int wl = INT_MAX, wh = -1;
        int q = lo;
        for (; q <= hi && (q & 31); q++) { const int x = ra[q].x; const int s0 = x & 0x0fffffff; wl = min(wl, s0); wh = max(wh, s0 + ((x >> 28) & 7) - 1); }
        for (; q + 31 <= hi; q += 32) { const int2 h = blk[q >> 5]; wl = min(wl, h.x); wh = max(wh, h.y); }
        for (; q <= hi; q++) { const int x = ra[q].x; const int s0 = x & 0x0fffffff; wl = min(wl, s0); wh = max(wh, s0 + ((x >> 28) & 7) - 1); }
        int* w = reinterpret_cast<int*>(&uwin[r.x & 0x0fffffff]);
        if (wl < w[0]) atomicMin(&w[0], wl);
        if (wh > w[1]) atomicMax(&w[1], wh);
    }
}

// out[dst] = out[src]   /   *total += sel[idx]   (tiny helpers of the incremental bookkeeping)
__global__ void k_add_selected(double* total, const double* v, int idx) { *total += v[idx]; }

// ------------------------------------------------------------------------------------------------
// repeat path: every pixel that touches a DUPLICATED data bin, evaluated per pixel exactly like the
// reference's pixel loop (kernels3.cu:2895-3220 == 3383-3700): float32 expected values summed over
// the ACTIVE copy pairs in the kernel's loop order, observed counts looked up in the contact lists.
// ------------------------------------------------------------------------------------------------
struct FragGeom { float mid[3]; int acc[3]; int lim, id_c, pos, circ, ori; float stot; };

__device__ __forceinline__ FragGeom frag_geom(const int* __restrict__ slot, int ld, int f, const LevelView& lv) {
    FragGeom G;
    const int id_d = slot[F_ID_D * ld + f];
    const int4 sid = lv.sub_id[id_d];
    G.lim = sid.w - 1;
    const float len[3] = { lv.sub_len[id_d * 3], lv.sub_len[id_d * 3 + 1], lv.sub_len[id_d * 3 + 2] };
    G.acc[0] = lv.sub_accu[id_d * 3]; G.acc[1] = lv.sub_accu[id_d * 3 + 1]; G.acc[2] = lv.sub_accu[id_d * 3 + 2];
    G.ori = slot[F_ORI * ld + f];
    G.id_c = slot[F_ID_C * ld + f]; G.pos = slot[F_POS * ld + f]; G.circ = slot[F_CIRC * ld + f];
    G.stot = __int2float_rn(slot[F_L_CONT_BP * ld + f]) / 1000.0f;
    const float start_kb = __int2float_rn(slot[F_START_BP * ld + f]) / 1000.0f;
    float accu = 0.0f;
    G.mid[0] = G.mid[1] = G.mid[2] = 0.0f;
    #pragma unroll
    for (int i = 0; i < 3; i++) {
        if (i > G.lim) break;
        const int loc = (G.ori == 1) ? i : G.lim - i;
        const float l = (loc == 0) ? len[0] : (loc == 1 ? len[1] : len[2]);
        float mid;
        if (i == 0) { mid = start_kb + l / 2.0f; accu = start_kb + l; }
        else { mid = accu + l / 2.0f; accu = accu + l; }
        if (loc == 0) G.mid[0] = mid; else if (loc == 1) G.mid[1] = mid; else G.mid[2] = mid;
    }
    return G;
}

__device__ __forceinline__ float lookup_obs(const long long* __restrict__ rowptr, const int2* __restrict__ contacts, int sa, int sb) {
    const int r = min(sa, sb), c = max(sa, sb);
    long long lo = rowptr[r], hi = rowptr[r + 1];
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if (contacts[mid].x < c) lo = mid + 1; else hi = mid; }
    return (lo < rowptr[r + 1] && contacts[lo].x == c) ? __int_as_float(contacts[lo].y) : 0.0f;
}

// evaluate_likelihood_double (kernels3.cu:191-210)
__device__ __forceinline__ double poisson_ll(float exf, float obf) {
    const double ex = (double)exf, ob = (double)obf;
    if (ex == 0.0) return 0.0;
    if (ob >= 15.0) return ob * log(ex) - ex - (ob * log(ob) - ob + log(sqrt(ob * 2.0 * M_PI)));
    if (ob > 0.0) return ob * log(ex) - ex - log((double)factorial_f32(obf));
    if (ob == 0.0) return -ex;
    return 0.0;
}

__device__ double repeat_pixel(const int* __restrict__ slot, int ld, const LevelView& lv, const int* __restrict__ collector,
                               const int2* __restrict__ dispatcher, const long long* __restrict__ rowptr,
                               const int2* __restrict__ contacts, int bi, int bj, bool diag, const Params& p) {
    float ex[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    const int2 di = dispatcher[bi], dj = dispatcher[bj];
    for (int ci = di.x; ci < di.y; ci++) {
        const int fi = collector[ci];
        if (slot[F_ACTIV * ld + fi] != 1) continue;
        const FragGeom Gi = frag_geom(slot, ld, fi, lv);
        for (int cj = dj.x; cj < dj.y; cj++) {
            const int fj = collector[cj];
            if (slot[F_ACTIV * ld + fj] != 1) continue;
            const FragGeom Gj = frag_geom(slot, ld, fj, lv);
            if (Gi.id_c == Gj.id_c) {
                const bool swap = Gi.pos > Gj.pos;              // the frag closest to the contig origin decides circ / s_tot
                const bool circ = (swap ? Gj.circ : Gi.circ) == 1;
                const float stot = swap ? Gj.stot : Gi.stot;
                #pragma unroll
                for (int a = 0; a < 3; a++) {
                    #pragma unroll
                    for (int b = 0; b < 3; b++) {
                        if (a > Gi.lim || b > Gj.lim) continue;
                        const float s = fabsf(Gj.mid[b] - Gi.mid[a]);
                        const float norm = __int2float_rn(Gi.acc[a] * Gj.acc[b]) / p.nfpb;
                        const float r = circ ? rippe_contacts_circ(s, stot, p) : rippe_contacts(s, p);
                        ex[a][b] = ex[a][b] + r * norm;
                    }
                }
            } else {
                #pragma unroll
                for (int a = 0; a < 3; a++) {
                    #pragma unroll
                    for (int b = 0; b < 3; b++) {
                        if (a > Gi.lim || b > Gj.lim) continue;
                        const int ai = (Gi.ori == 1) ? Gi.acc[a] : ((Gi.lim == 0) ? Gi.acc[0] : (Gi.lim == 1 ? Gi.acc[1] : Gi.acc[2]));   // quirk Q1
                        const float norm = __int2float_rn(ai * Gj.acc[b]) / p.nfpb;
                        ex[a][b] = ex[a][b] + p.v_inter * norm;
                    }
                }
            }
        }
    }
    const int4 si = lv.sub_id[bi], sj = lv.sub_id[bj];
    double ll = 0.0;
    #pragma unroll
    for (int a = 0; a < 3; a++) {
        #pragma unroll
        for (int b = 0; b < 3; b++) {
            if (a >= si.w || b >= sj.w || (diag && b <= a)) continue;
            ll += poisson_ll(ex[a][b], lookup_obs(rowptr, contacts, si.x + a, sj.x + b));
        }
    }
    return ll;
}

// mark the duplicated data bins that have a copy inside U (sub_index_repeats of cuda_lib_gl.py:2458)
__global__ void k_mark_rep_in_u(const int* __restrict__ base, int ld, const int* __restrict__ sub_index, const int* __restrict__ meta,
                                LevelView lv, unsigned char* __restrict__ rep_in_u) {
    const int m = meta[4];
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < m; u += gridDim.x * blockDim.x) {
        const int d = base[F_ID_D * ld + sub_index[u]];
        if (lv.dup[d]) rep_in_u[d] = 1;
    }
}

// Pixel enumeration t in [0, n_rep * N): r = rep_bins[t / N], j = t % N.
//   FULL : every pixel touching a duplicated bin once (j unique, j == r -> diagonal, j duplicated and j > r)
//   DELTA: ranges 2-4 of sub_compute_likelihood (kernels3.cu:3363-3380): r must have a copy in U;
//          r x every unique bin, r x r' for duplicated r' > r also in U, and the diagonal of r
template <bool DELTA>
__global__ void __launch_bounds__(128)
k_repeat_pixels(const int* __restrict__ slot0, size_t slot_stride, int ld, LevelView lv, const int* __restrict__ collector,
                const int2* __restrict__ dispatcher, const long long* __restrict__ rowptr, const int2* __restrict__ contacts,
                const int* __restrict__ rep_bins, int n_rep, const unsigned char* __restrict__ rep_in_u,
                const __grid_constant__ Params p, double* __restrict__ partials, int partial_stride) {
    const int k = blockIdx.y;
    const int* slot = slot0 + (size_t)k * slot_stride;
    const int N = lv.n_data;
    double acc = 0.0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < (long long)n_rep * N; t += (long long)gridDim.x * blockDim.x) {
        const int r = rep_bins[t / N], j = (int)(t % N);
        if (DELTA && !rep_in_u[r]) continue;
        const bool jdup = lv.dup[j] != 0;
        if (jdup && j < r) continue;
        if (DELTA && jdup && j != r && !rep_in_u[j]) continue;
        acc += repeat_pixel(slot, ld, lv, collector, dispatcher, rowptr, contacts, min(r, j), max(r, j), j == r, p);
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) partials[(size_t)k * partial_stride + blockIdx.x] = acc;
}

// ------------------------------------------------------------------------------------------------
// statistics and the distance histogram
// ------------------------------------------------------------------------------------------------
__global__ void k_stats(const int* __restrict__ slot, int ld, int n, unsigned long long* __restrict__ acc) {
    // acc[0] = #heads (start_bp == 0), acc[1] = sum l_cont_bp over heads, acc[2] = min l_cont, acc[3] = max l_cont
    unsigned long long heads = 0, sum = 0; int mn = INT_MAX, mx = INT_MIN;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int lc = slot[F_L_CONT * ld + i];
        mn = min(mn, lc); mx = max(mx, lc);
        if (slot[F_START_BP * ld + i] == 0) { heads++; sum += (unsigned long long)(long long)slot[F_L_CONT_BP * ld + i]; }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {                      // one set of atomics per warp, not per thread
        heads += __shfl_down_sync(0xffffffffu, heads, o); sum += __shfl_down_sync(0xffffffffu, sum, o);
        mn = min(mn, __shfl_down_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (heads) { atomicAdd(&acc[0], heads); atomicAdd(&acc[1], sum); }
        atomicMin((int*)&acc[2], mn); atomicMax((int*)&acc[3], mx);
    }
}
__global__ void k_stats_final(const unsigned long long* __restrict__ acc, const int* __restrict__ n_contigs, double* __restrict__ out) {
    out[0] = (double)*n_contigs;
    out[1] = (double)*(const int*)&acc[2];
    out[2] = acc[0] ? (double)acc[1] / (double)acc[0] : 0.0;
    out[3] = (double)*(const int*)&acc[3];
}
__global__ void k_count_contigs(const int* __restrict__ first, int cap, int* __restrict__ n_contigs) {
    int c = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) c += first[i] != INT_MAX;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(n_contigs, c);
}

// dist_inter_genome (cuda_lib_gl.py:475-541): per-bin neighbour / orientation agreement with the initial
// genome; every term is a multiple of 0.5, so the float64 sum is exact whatever the order.
__global__ void __launch_bounds__(256)
k_dist_genome(const int* __restrict__ slot0, size_t slot_stride, int ld, int n, const int* __restrict__ init_prev, const int* __restrict__ init_next,
              const int* __restrict__ orientable, const unsigned char* __restrict__ skip, double* __restrict__ partials, int partial_stride) {
    const int* slot = slot0 + (size_t)blockIdx.y * slot_stride;          // blockIdx.y: one of several consecutive slots
    double acc = 0.0;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < n; f += gridDim.x * blockDim.x) {
        if (skip[f]) continue;
        const int p0 = init_prev[f], n0 = init_next[f];
        const int tp = slot[F_PREV * ld + f], tn = slot[F_NEXT * ld + f];
        int p1 = (tp != -1) ? slot[F_ID_D * ld + tp] : tp;
        int n1 = (tn != -1) ? slot[F_ID_D * ld + tn] : tn;
        double d = 3.0;
        if ((p1 == p0 && n1 == n0) || (p1 == n0 && n1 == p0)) d -= 1.0;
        if (orientable[f]) {
            int swap = 1;
            if (slot[F_ORI * ld + f] != 1) { const int t = p1; p1 = n1; n1 = t; swap = -1; }     // initial orientation is +1
            if (p0 == p1) {
                if (p0 == -1) d -= 1.0;
                else if (!orientable[p1]) d -= 1.0;
                else { d -= 0.5; if (1 == swap * slot[F_ORI * ld + p1]) d -= 0.5; }
            }
            if (n0 == n1) {
                if (n0 == -1) d -= 1.0;
                else if (!orientable[n1]) d -= 1.0;
                else { d -= 0.5; if (1 == swap * slot[F_ORI * ld + n1]) d -= 0.5; }
            }
        } else {
            if (p1 == p0 || p1 == n0) d -= 1.0;
            if (n1 == n0 || n1 == p0) d -= 1.0;
        }
        acc += d;
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) partials[(size_t)blockIdx.y * partial_stride + blockIdx.x] = acc;
}

// distance histogram of estimate_parameters (cuda_lib_gl.py:1236-1270).  The initial sub-level layout
// is contig-contiguous and position-ordered, so the cis partners of sub-frag i are i+1 .. end of contig.
// One warp per sub-frag i: pair counts per bin over j > i (zeros included), contacts from row i.
__global__ void __launch_bounds__(256)
k_dist_hist(const long long* __restrict__ rowptr, const int2* __restrict__ contacts, int W,
            const int* __restrict__ id_c, const int* __restrict__ st, const int* __restrict__ ln, const int* __restrict__ pos,
            double max_kb, double bin_kb, int n_bins, double* __restrict__ d_sum, unsigned long long* __restrict__ d_cnt) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < W - 1; i += n_warps) {
        const int ci = id_c[i], sti = st[i], lni = ln[i], pi = pos[i];
        for (int base = i + 1; base < W; base += 32) {
            const int j = base + lane;
            bool live = j < W && id_c[j] == ci;
            if (live) {
                double d = (pi < pos[j]) ? ((double)(st[j] - sti - lni) + (double)(lni + ln[j]) / 2.) / 1000.
                                         : ((double)(sti - st[j] - ln[j]) + (double)(ln[j] + lni) / 2.) / 1000.;
                if (d < max_kb) { const int b = (int)(d / bin_kb); if (b >= 0 && b < n_bins) atomicAdd(&d_cnt[b], 1ull); }
            }
            if (!__any_sync(0xffffffffu, live)) break;      // contig-contiguous layout
        }
        for (long long e = rowptr[i] + lane; e < rowptr[i + 1]; e += 32) {
            const int2 ce = contacts[e];
            const int j = ce.x;
            if (id_c[j] != ci) continue;
            double d = (pi < pos[j]) ? ((double)(st[j] - sti - lni) + (double)(lni + ln[j]) / 2.) / 1000.
                                     : ((double)(sti - st[j] - ln[j]) + (double)(ln[j] + lni) / 2.) / 1000.;
            if (d < max_kb) { const int b = (int)(d / bin_kb); if (b >= 0 && b < n_bins) atomicAdd(&d_sum[b], (double)__int_as_float(ce.y)); }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
// per-kernel device timers (CUDA event pairs on the context's stream), enabled on demand
#define PROF_KERNELS 8
#define PROF_RING 4096
struct Profiler {
    bool on = false;
    std::vector<cudaEvent_t> ev[PROF_KERNELS];      // 2 * PROF_RING events per kernel id, created lazily
    int head[PROF_KERNELS] = {0}, pending[PROF_KERNELS] = {0};
    double total_ms[PROF_KERNELS] = {0};
    long long count[PROF_KERNELS] = {0};
    void resolve(int k) {
        for (int i = 0; i < pending[k]; i++) {
            const int slot = (head[k] - pending[k] + i + PROF_RING) % PROF_RING;
            float ms = 0.f;
            cudaEventSynchronize(ev[k][2 * slot + 1]);
            if (cudaEventElapsedTime(&ms, ev[k][2 * slot], ev[k][2 * slot + 1]) == cudaSuccess) { total_ms[k] += ms; count[k]++; }
        }
        pending[k] = 0;
    }
    void begin(int k, cudaStream_t st) {
        if (!on) return;
        if (ev[k].empty()) { ev[k].resize(2 * PROF_RING); for (auto& e : ev[k]) cudaEventCreate(&e); }
        if (pending[k] == PROF_RING) resolve(k);
        cudaEventRecord(ev[k][2 * head[k]], st);
    }
    void end(int k, cudaStream_t st) {
        if (!on) return;
        cudaEventRecord(ev[k][2 * head[k] + 1], st);
        head[k] = (head[k] + 1) % PROF_RING; pending[k]++;
    }
    void destroy() { for (int k = 0; k < PROF_KERNELS; k++) { for (auto& e : ev[k]) cudaEventDestroy(e); ev[k].clear(); } }
};

// Scratch of one proposal being scored.  Proposals of one step are independent given the base slot, so
// graal_score_proposal runs each on its own lane (stream + buffers) and their kernels overlap on the GPU;
// every other entry point joins the lanes back into the context stream first.
#define GRAAL_MAX_LANES 4
struct Lane {
    cudaStream_t st = nullptr; cudaEvent_t done = nullptr; bool pending = false;
    // side streams of the lane: the contact pass, the candidate band pass and the base band pass of a proposal only
    // depend on the geometry / order kernels, not on each other (fork / join by events; graph branches when captured)
    cudaStream_t side[2] = {nullptr, nullptr}; cudaEvent_t ev_ready = nullptr, ev_side[2] = {nullptr, nullptr};
    int cand_first = -1;                     // candidate slots [cand_first, cand_first + 13) in use while pending
    int* ints = nullptr;                     // [256]: [8..16) delta meta, [16..120) piece_len, [160..188) changed ranges
    int* sub_index = nullptr;                // [n]
    unsigned* chmask = nullptr;              // [W] bit k: record differs from the base slot in candidate k
    unsigned* chmask2 = nullptr;             // [W] bit k (paired candidates): record differs from candidate k - 1
    Geo* geo_cand = nullptr;                 // [13][W]
    int4* cand_ordrec = nullptr; int4* base_ordrec = nullptr;   // [13][n] / [n] position-ordered bin records A of U
    int4* cand_ordb = nullptr; int4* base_ordb = nullptr; int4* base_ordc = nullptr;   // records B (mid-points) and C (masks)
    double* partials = nullptr;              // [48][partial_stride]: contacts rows 0..12, candidate band 16..28, base band 32..44
    double* dist_partials = nullptr;         // [13][partial_stride]: genome distance of the candidates
    unsigned char* rep_in_u = nullptr;       // [N]
    int2* uwin = nullptr;                    // [W] union window of the bins of U (first sub-frag of the bin)
    int2* cand_blk = nullptr;                // [13][n / 32 + 1] block hulls of the candidate orders
};

// The ~20 launches that score one proposal, captured once per proposal index and replayed with the two
// kernels that take (id_fA, id_fB, max_id) by value re-parameterised: one graph launch instead of twenty
// kernel launches on the host's critical path between two draws.
struct ProposalGraph {
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    cudaGraphNode_t n_build = nullptr, n_setup = nullptr;
    cudaKernelNodeParams kp_build{}, kp_setup{};
    int base_slot = -1, first_cand = -1; double* d_out = nullptr; double* d_dist = nullptr; unsigned skip = 0u; long long version = -1; cudaStream_t st = nullptr;
    int n_launches = 0;
    void reset() {
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        exec = nullptr; graph = nullptr; n_build = n_setup = nullptr; version = -1;
    }
};

// A launch sequence whose arguments do not change from call to call (state statistics, relabel): captured
// once per argument set, replayed with one graph launch.
struct FixedGraph {
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    long long version = -1, k0 = 0, k1 = 0, k2 = 0; int n_launches = 0;
    void reset() {
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        exec = nullptr; graph = nullptr; version = -1;
    }
};

struct graal_ctx {
    Profiler prof;
    FixedGraph g_stats, g_relabel, g_full, g_full_cached;
    Lane lanes[GRAAL_MAX_LANES]; int n_lanes = 3; cudaEvent_t ev_fork = nullptr, ev_fetch = nullptr; bool fetch_pending = false, fetch_ncontigs = false;
    int fetch_seq = 0; bool fetch_published = false, publish_enable = true;     // results written by a kernel into mapped pinned memory + a sequence word the host polls
    long long commits_total = 0, fetch_mark = 0;                 // commits on the bound slot, and their count when the pending readback left
    int fork_passes = 1;                     // GRAAL_FORK=0: contact / band passes of a proposal on one stream
    int pairing = 1;                         // GRAAL_PAIRING=0: score candidates 3, 5, 7 like the others (A/B runs)
    ProposalGraph graphs[16]; long long version = 0; int use_graphs = 1;      // version: bumped whenever captured arguments go stale
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int n_sm = 148;
    int64_t launches = 0;
    // level
    int N = 0, n_new = 0, W = 0;
    long long E = 0;
    LevelView lv{};
    const int* collector = nullptr; const int* dispatcher = nullptr;
    const long long* rowptr = nullptr; const int2* contacts = nullptr;
    float nfpb = 1.0f;
    std::vector<std::pair<int, long long>> accu_hist;   // (accu value, count) over all sub-frags; index = position
    unsigned char* d_accu_idx = nullptr;                // [N*3]
    float* d_tab_norm = nullptr; float* d_tab_g[2] = {nullptr, nullptr}; double* d_tab_logg[2] = {nullptr, nullptr};
    double* d_tab_lnnorm = nullptr; double2* d_tab_log = nullptr; double* d_tab_exp = nullptr;
    int4* d_tab_lnf[2] = {nullptr, nullptr}; int4* d_tab_f[2] = {nullptr, nullptr}; int4* d_tab_lnfu[2] = {nullptr, nullptr}; int4* d_tab_lnfu7[2] = {nullptr, nullptr}; int4* d_tab_fu[2] = {nullptr, nullptr}; double* d_tab_normd = nullptr;
    std::vector<double> h_law, h_law7;
    int math_mode = 2;
    std::vector<float> h_tab_g; std::vector<double> h_tab_logg;
    int* d_quirky = nullptr; int n_quirky = 0;
    unsigned char* d_dup = nullptr; unsigned char* d_sub_dup = nullptr;
    int* d_rep_bins = nullptr; int n_rep = 0;
    double lf_total = 0.0, ob_total = 0.0;
    unsigned short* cid16_base = nullptr; float* mid32_base = nullptr; int2* cm_base = nullptr; int smem_optin = 0;
    int* group_row = nullptr; int n_groups = 0;
    // windowed contact pass: static row items, per-pass order-space arrays and row records
    int4* items = nullptr; int n_items = 0; int4* row_hdr = nullptr; int4* item_hdr = nullptr;
    int* o_start = nullptr; int2* o_sub = nullptr; int* o_cid = nullptr; int2* o_blk = nullptr;
    int band_split = 1;                       // GRAAL_BAND_SPLIT=1..32: most warps the band of one bin of a short U is cut over in the fast band delta kernels
    int delta_minb = 2;                       // GRAAL_DELTA_MINB=2|3|4: resident CTAs per SM the delta contact pass is compiled for (3, 4: accumulators in shared memory)
    int delta_split = DC_MAX_SPLIT;           // GRAAL_DELTA_SPLIT=1|2|4|8: most slices a row of a short U is cut into in the delta contact pass
    int delta_rel = 1;                        // GRAAL_DELTA_REL=0: absolute contact terms in the delta contact pass on every level (A/B runs)
    int band_fast = 1;                        // GRAAL_BAND_FAST=0: general band delta kernels on every level (A/B runs)
    int delta_uni = 0;                        // GRAAL_DELTA_UNI=1: union windows in the delta contact pass (measured: no gain -- the pass is bound by the evaluation of the in-band entries)
    float max_bin_kb = 0.0f;                  // longest bin of the level (reach of the candidate windows)
    int win_slot = -1; float win_dmax = 0.0f; long long geo_epoch = 0, win_epoch = -1;   // slot / d_max / geometry the row records (row_hdr) describe
    FixedGraph g_win;
    int full_win = 1;                         // GRAAL_FULL_WIN=0: gather-everything kernel (k_full_contacts_direct) for A/B runs
    int win_stab = 0; bool win_stab_allowed = true;   // law table of the windowed full pass in shared memory (coarse copy); GRAAL_WIN_STAB=0: never, =1: always (no timing)
    bool win_tuned = false;                   // SUB picked by timing the variants on the bound level (first windowed pass)
    int win_unroll = 8, win_minb = 4, win_sub = 2;   // GRAAL_WIN_SUB=1|2|4|8: loads per exact-path group; GRAAL_WIN_UNROLL=2|4|8: stream loads in flight per warp; GRAAL_WIN_MINB=2..6: CTAs per SM
    int smem_cid = 0;                         // GRAAL_SMEM_CID=1: stage the contig-id table in shared memory (measured: no faster than L1, profiles/README.md)
    double* band_hist = nullptr;              // [16][13] band delta of the proposals scored since the last commit
    int band_slot = -1; int band_age = 0;     // slot whose cross-bin band total is cached in d_scalars[40]
    // params
    Params p{}; bool have_params = false;
    // state
    int* slots = nullptr; int ld = 0, n_slots = 0;
    // scratch
    Geo* geo_base = nullptr; int geo_base_slot = -1;
    int first_idx_slot = -1;                 // slot whose first-bin-of-contig table (first_idx) is current
    // fused prologue: host-side bound on the contig ids of `bound_slot` (n_contigs after the last relabel, read back with the
    // step's fetch, + the ids one committed candidate can add); -1: unknown -> the general path
    int contig_bound = -1, bound_slot = -1, bound_commits = 0; int* h_ncontigs = nullptr; bool ncontigs_pending = false;
    bool first_clean = false, stats_clean = false; FixedGraph g_prologue; int fused_prologue = 1;
    int* order = nullptr;                    // [n]
    int* cont_len = nullptr; int* cont_off = nullptr; int cap = 0;   // [cap]
    int* first_idx = nullptr; int* map = nullptr;
    unsigned long long* keys = nullptr; unsigned long long* keys_sorted = nullptr;
    void* cub_tmp = nullptr; size_t cub_tmp_bytes = 0; int key_bits = 32;
    int* d_ints = nullptr;                   // [0]=max_id [1]=n_contigs [2]=err [3]=tmp max  [8..16)=delta meta  [16..16+13*8)=piece_len
    unsigned long long* d_stats = nullptr;   // [4]
    unsigned long long* d_counters = nullptr; // [4] profiler counters (k_u_stats)
    double* partials = nullptr; int partial_stride = 0;   // [14][partial_stride]
    double* d_scalars = nullptr;             // [16] device doubles: [0]=contacts [1]=band [2]=quirk ...
};

static inline int nblk(long long n, int b) { return (int)((n + b - 1) / b); }
static size_t slot_stride(const graal_ctx* c) { return (size_t)N_FIELDS * c->ld; }
static int* slot_ptr(const graal_ctx* c, int s) { return c->slots + (size_t)s * slot_stride(c); }

static double host_g0(const graal_ctx* c, const Params& p) {
    // sum over all unordered sub-frag pairs of g(a*b), grouped by the distinct accu values
    double tot = 0.0;
    const auto& h = c->accu_hist;
    for (size_t i = 0; i < h.size(); i++) {
        const double ci = (double)h[i].second;
        tot += ci * (ci - 1.0) / 2.0 * (double)g_clamp(h[i].first * h[i].first, p.v_inter, p.nfpb);
        for (size_t j = i + 1; j < h.size(); j++)
            tot += ci * (double)h[j].second * (double)g_clamp(h[i].first * h[j].first, p.v_inter, p.nfpb);
    }
    return tot;
}

// Fill the nd x nd clamp tables of a parameter set into table slot `which` (0: current parameters,
// 1: test parameters of the nuisance step) and point `p` at them.
static int upload_tables(graal_ctx* c, Params& p, int which) {
    const int nd = (int)c->accu_hist.size();
    c->h_tab_g.resize((size_t)nd * nd); c->h_tab_logg.resize((size_t)nd * nd);
    for (int i = 0; i < nd; i++) for (int j = 0; j < nd; j++) {
        const float g = g_clamp(c->accu_hist[i].first * c->accu_hist[j].first, p.v_inter, p.nfpb);
        c->h_tab_g[(size_t)i * nd + j] = g;
        c->h_tab_logg[(size_t)i * nd + j] = (g != 0.0f) ? log((double)g) : -(double)INFINITY;   // g < 0 -> NaN, as log(ex) in the reference
    }
    // pageable -> device: the runtime stages the buffer before returning, so the host vectors can be reused
    CUDA_OK(cudaMemcpyAsync(c->d_tab_g[which], c->h_tab_g.data(), (size_t)nd * nd * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(c->d_tab_logg[which], c->h_tab_logg.data(), (size_t)nd * nd * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    p.nd = nd; p.t_norm = c->d_tab_norm; p.t_g = c->d_tab_g[which]; p.t_logg = c->d_tab_logg[which];
    p.mode = c->math_mode; p.t_lnnorm = c->d_tab_lnnorm; p.t_log = c->d_tab_log; p.t_exp = c->d_tab_exp;
    const double cf = (double)p.c1 * (double)p.fact;
    p.ln_cf = cf > 0.0 ? log(cf) : -(double)INFINITY;           // c1*fact <= 0: rippe <= 0, the clamp wins
    p.ln_v = p.v_inter > 0.0f ? log((double)p.v_inter) : (p.v_inter == 0.0f ? -(double)INFINITY : (double)NAN);
    p.slope_d = (double)p.slope;
    // tabulated law: f(s) = c1*fact * s^slope * exp((d-2)/(x^2+d)), x = s*lm/kuhn, and ln f, with d/ds, float64
    p.t_lnf = c->d_tab_lnf[which]; p.t_f = c->d_tab_f[which]; p.t_normd = c->d_tab_normd; p.v_clamp = (double)p.v_inter;
    p.t_lnfu = c->d_tab_lnfu[which]; p.t_fu = c->d_tab_fu[which]; p.fu_ok = 0; p.fu_span = p.fu_zlo = p.fu_zspan = 0u;
    p.t_lnfu7 = c->d_tab_lnfu7[which]; p.fu7_smin = 0u; p.fu7_span = 0u;
    if (p.mode == 2) {
        if (!(cf > 0.0) || !(p.kuhn > 0.0f) || !(p.lm > 0.0f)) p.mode = 1;      // degenerate parameters: analytic path
        else {
            struct Entry { double a0; float a1, a2; };
            static_assert(sizeof(Entry) == 16, "law table entry");
            const size_t half = (size_t)(LAW_NODES + 2) * 2;          // doubles per table (16 bytes per interval)
            c->h_law.assign(half * 4, 0.0);
            Entry* e_ln = reinterpret_cast<Entry*>(c->h_law.data());
            Entry* e_f = reinterpret_cast<Entry*>(c->h_law.data() + half);
            const double K = (double)p.d - 2.0, dd = (double)p.d, q = (double)p.lm / (double)p.kuhn, sl = (double)p.slope;
            auto lnf = [&](double sv) { const double x = sv * q; return p.ln_cf + sl * log(sv) + K / (x * x + dd); };
            // quadratic in u = 1 + t through the Chebyshev nodes t_j = 1/2 + cos((2j+1) pi/6)/2 of the interval
            double un[3], w[3][3];
            for (int j = 0; j < 3; j++) un[j] = 1.5 + 0.5 * cos((2 * j + 1) * 3.14159265358979323846 / 6.0);
            for (int j = 0; j < 3; j++) {         // Lagrange basis of node j expanded in powers of u
                const double u1 = un[(j + 1) % 3], u2 = un[(j + 2) % 3], den = (un[j] - u1) * (un[j] - u2);
                w[j][0] = u1 * u2 / den; w[j][1] = -(u1 + u2) / den; w[j][2] = 1.0 / den;
            }
            for (int i = 0; i < LAW_NODES; i++) {
                const int e = LAW_EMIN + (i >> LAW_M);
                const double lo = ldexp(1.0 + (double)(i & ((1 << LAW_M) - 1)) / (double)(1 << LAW_M), e), h = ldexp(1.0, e - LAW_M);
                double al[3] = {0, 0, 0}, af[3] = {0, 0, 0};
                for (int j = 0; j < 3; j++) {
                    const double v = lnf(lo + h * (un[j] - 1.0)), f = exp(v);
                    for (int k = 0; k < 3; k++) { al[k] += v * w[j][k]; af[k] += f * w[j][k]; }
                }
                e_ln[i].a0 = al[0]; e_ln[i].a1 = (float)al[1]; e_ln[i].a2 = (float)al[2];
                e_f[i].a0 = af[0]; e_f[i].a1 = (float)af[1]; e_f[i].a2 = (float)af[2];
            }
            CUDA_OK(cudaMemcpyAsync(c->d_tab_lnf[which], c->h_law.data(), half * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            CUDA_OK(cudaMemcpyAsync(c->d_tab_f[which], c->h_law.data() + half, half * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            // Uniform-accu level (one accu value, finite positive clamp): the windowed contact pass adds
            // ob * (ln f(s) + ln norm - log g) for the in-band entries.  With a law that DECREASES over the table range
            // the clamp max(f, v_inter) is a threshold on s: beyond s_c the pixel evaluates to exactly g and contributes 0.
            const float g1 = (nd == 1) ? g_clamp(c->accu_hist[0].first * c->accu_hist[0].first, p.v_inter, p.nfpb) : 0.0f;
            if (nd == 1 && g1 > 0.0f && p.d_max > 0.0f && p.v_inter > 0.0f) {
                const float tn = (float)(c->accu_hist[0].first * c->accu_hist[0].first) / p.nfpb;
                const double cst = log((double)tn) - log((double)g1);
                auto f2b = [](float f) { unsigned u; memcpy(&u, &f, 4); return u; };
                auto b2f = [](unsigned u) { float f; memcpy(&f, &u, 4); return f; };
                const unsigned b_lo = LAW_SMIN_BITS, b_tab = f2b(ldexpf(1.0f, LAW_EMAX)), b_dmax = f2b(p.d_max);
                bool mono = cst - cst == 0.0;
                double prev = lnf((double)b2f(b_lo));
                for (int i = 1; i <= LAW_NODES && mono; i++) {        // decreasing at every interval edge
                    const double v = lnf(ldexp(1.0 + (double)(i & ((1 << LAW_M) - 1)) / (double)(1 << LAW_M), LAW_EMIN + (i >> LAW_M)));
                    if (!(v < prev)) mono = false;
                    prev = v;
                }
                if (mono) {
                    unsigned lo = b_lo, hi = b_tab;                  // smallest pattern in [b_lo, b_tab] with ln f <= ln v (b_tab: none inside the table)
                    if (lnf((double)b2f(lo)) > p.ln_v) {
                        while (hi - lo > 1u) { const unsigned m = lo + (hi - lo) / 2u; if (lnf((double)b2f(m)) > p.ln_v) lo = m; else hi = m; }
                    } else hi = lo;
                    const unsigned b_sc = (hi >= b_tab && lnf((double)b2f(b_tab)) > p.ln_v) ? 0x7f800000u : hi;
                    const unsigned b_hi = std::min(std::min(b_dmax, b_tab), b_sc);
                    p.fu_span = b_hi > b_lo ? b_hi - b_lo : 0u;
                    const unsigned z_hi = std::min(b_dmax, b_sc);
                    p.fu_zlo = b_tab; p.fu_zspan = z_hi > b_tab ? z_hi - b_tab : 0u;
                    Entry* e_u = reinterpret_cast<Entry*>(c->h_law.data() + 2 * half);
                    for (int i = 0; i < LAW_NODES; i++) { e_u[i] = e_ln[i]; e_u[i].a0 += cst; }
                    CUDA_OK(cudaMemcpyAsync(c->d_tab_lnfu[which], c->h_law.data() + 2 * half, half * sizeof(double), cudaMemcpyHostToDevice, c->stream));
                    // coarse copy for shared memory: the LAW7_OCT octaves that end where the fast range ends (shorter in-band
                    // distances, if any, take the exact evaluation)
                    if (p.fu_span > 0u) {
                        const int e_top = (int)((b_hi - 1u) >> 23) - 127;                 // octave of the last fast distance
                        const int e_lo7 = std::max(LAW_EMIN, std::min(e_top - LAW7_OCT + 1, LAW_EMAX - LAW7_OCT));
                        c->h_law7.resize((size_t)LAW7_NODES * 2);
                        Entry* e7 = reinterpret_cast<Entry*>(c->h_law7.data());
                        for (int i = 0; i < LAW7_NODES; i++) {
                            const int e = e_lo7 + (i >> LAW7_M);
                            const double lo7 = ldexp(1.0 + (double)(i & ((1 << LAW7_M) - 1)) / (double)(1 << LAW7_M), e), h7 = ldexp(1.0, e - LAW7_M);
                            double al[3] = {0, 0, 0};
                            for (int j = 0; j < 3; j++) { const double v = lnf(lo7 + h7 * (un[j] - 1.0)); for (int k = 0; k < 3; k++) al[k] += v * w[j][k]; }
                            e7[i].a0 = al[0] + cst; e7[i].a1 = (float)al[1]; e7[i].a2 = (float)al[2];
                        }
                        CUDA_OK(cudaMemcpyAsync(c->d_tab_lnfu7[which], c->h_law7.data(), (size_t)LAW7_NODES * sizeof(int4), cudaMemcpyHostToDevice, c->stream));
                        p.fu7_smin = (unsigned)(127 + e_lo7) << 23;
                        const unsigned b_top7 = (unsigned)(127 + e_lo7 + LAW7_OCT) << 23;
                        const unsigned b_hi7 = std::min(b_hi, b_top7);
                        p.fu7_span = b_hi7 > p.fu7_smin ? b_hi7 - p.fu7_smin : 0u;
                    }
                    Entry* e_fu = reinterpret_cast<Entry*>(c->h_law.data() + 3 * half);
                    for (int i = 0; i < LAW_NODES; i++) {
                        e_fu[i].a0 = e_f[i].a0 * (double)tn - (double)g1;
                        e_fu[i].a1 = (float)((double)e_f[i].a1 * (double)tn); e_fu[i].a2 = (float)((double)e_f[i].a2 * (double)tn);
                    }
                    CUDA_OK(cudaMemcpyAsync(c->d_tab_fu[which], c->h_law.data() + 3 * half, half * sizeof(double), cudaMemcpyHostToDevice, c->stream));
                    p.fu_ok = 1;
                }
            }
            CUDA_OK(cudaStreamSynchronize(c->stream));      // h_law is reused by the next call
        }
    }
    return 0;
}

// make the context stream wait for every proposal still running on a lane
static int join_lanes(graal_ctx* c) {
    for (int l = 0; l < c->n_lanes; l++) {
        Lane& L = c->lanes[l];
        if (!L.pending) continue;
        CUDA_OK(cudaStreamWaitEvent(c->stream, L.done, 0));
        L.pending = false; L.cand_first = -1;
    }
    return 0;
}

template <class F>
static int run_graphed(graal_ctx* c, FixedGraph& G, long long k0, long long k1, long long k2, F&& enqueue) {
    if (!c->use_graphs || c->prof.on) return enqueue();
    if (!(G.exec && G.version == c->version && G.k0 == k0 && G.k1 == k1 && G.k2 == k2)) {
        G.reset();
        const int64_t before = c->launches;
        CUDA_OK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue();
        cudaGraph_t graph = nullptr;
        const cudaError_t e_end = cudaStreamEndCapture(c->stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e_end != cudaSuccess) return set_err(-2, "stream capture failed: %s", cudaGetErrorString(e_end));
        G.graph = graph; G.n_launches = (int)(c->launches - before); c->launches = before;
        CUDA_OK(cudaGraphInstantiate(&G.exec, graph, 0));
        G.version = c->version; G.k0 = k0; G.k1 = k1; G.k2 = k2;
    }
    CUDA_OK(cudaGraphLaunch(G.exec, c->stream));
    c->launches += G.n_launches;
    return 0;
}

extern "C" {

const char* graal_last_error(void) { return g_err; }
const char* graal_version(void) { return GRAAL_VERSION; }

int graal_ctx_create(int device, graal_ctx** out) {
    if (!out) return set_err(-1, "null out pointer");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev <= 0) return set_err(-2, "no CUDA device available (%s): this library has no CPU path", cudaGetErrorString(e));
    if (device < 0 || device >= n_dev) return set_err(-1, "device %d out of range (0..%d)", device, n_dev - 1);
    CUDA_OK(cudaSetDevice(device));
    graal_ctx* c = new graal_ctx();
    c->device = device;
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    c->n_sm = prop.multiProcessorCount;
    c->smem_optin = (int)prop.sharedMemPerBlockOptin;
    { const char* e = getenv("GRAAL_SMEM_CID"); if (e && e[0] == '1') c->smem_cid = 1; }
    CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    CUDA_OK(cudaMalloc(&c->d_ints, 256 * sizeof(int)));
    CUDA_OK(cudaMemset(c->d_ints, 0, 256 * sizeof(int)));
    CUDA_OK(cudaMalloc(&c->d_stats, 4 * sizeof(unsigned long long)));
    CUDA_OK(cudaMalloc(&c->d_counters, 4 * sizeof(unsigned long long)));
    CUDA_OK(cudaMemset(c->d_counters, 0, 4 * sizeof(unsigned long long)));
    CUDA_OK(cudaMalloc(&c->d_scalars, 64 * sizeof(double)));
    c->partial_stride = c->n_sm * 8;
    CUDA_OK(cudaMalloc(&c->partials, (size_t)16 * c->partial_stride * sizeof(double)));
    { const char* e = getenv("GRAAL_GRAPHS"); if (e && e[0] == '0') c->use_graphs = 0; }
    { const char* e = getenv("GRAAL_PAIRING"); if (e && e[0] == '0') c->pairing = 0; }
    { const char* e = getenv("GRAAL_FORK"); if (e && e[0] == '0') c->fork_passes = 0; }
    { const char* e = getenv("GRAAL_FULL_WIN"); if (e && e[0] == '0') c->full_win = 0; }
    { const char* e = getenv("GRAAL_FUSED_PROLOGUE"); if (e && e[0] == '0') c->fused_prologue = 0; }
    CUDA_OK(cudaHostAlloc(&c->h_ncontigs, 64, cudaHostAllocMapped));      // [0] n_contigs readback, [8] sequence number of the last published fetch
    c->h_ncontigs[0] = 0; c->h_ncontigs[8] = 0;
    { const char* e = getenv("GRAAL_DELTA_UNI"); if (e) c->delta_uni = e[0] == '1'; }
    { const char* e = getenv("GRAAL_BAND_FAST"); if (e && e[0] == '0') c->band_fast = 0; }
    { const char* e = getenv("GRAAL_DELTA_REL"); if (e && e[0] == '0') c->delta_rel = 0; }
    { const char* e = getenv("GRAAL_BAND_SPLIT"); if (e && atoi(e) >= 1 && atoi(e) <= 32) c->band_split = atoi(e); }
    { const char* e = getenv("GRAAL_PUBLISH"); if (e && e[0] == '0') c->publish_enable = false; }
    { const char* e = getenv("GRAAL_DELTA_MINB"); if (e && atoi(e) >= 2 && atoi(e) <= 4) c->delta_minb = atoi(e); }
    { const char* e = getenv("GRAAL_DELTA_SPLIT"); if (e && atoi(e) >= 1 && atoi(e) <= 32) c->delta_split = atoi(e); }
    { const char* e = getenv("GRAAL_WIN_UNROLL"); if (e && (e[0] == '2' || e[0] == '4' || e[0] == '8')) c->win_unroll = e[0] - '0'; }
    { const char* e = getenv("GRAAL_WIN_MINB"); if (e && e[0] >= '2' && e[0] <= '6') c->win_minb = e[0] - '0'; }
    { const char* e = getenv("GRAAL_WIN_SUB"); if (e && (e[0] == '1' || e[0] == '2' || e[0] == '4' || e[0] == '8')) { c->win_sub = e[0] - '0'; c->win_tuned = true; } }
    { const char* e = getenv("GRAAL_WIN_STAB"); if (e && e[0] == '0') c->win_stab_allowed = false; if (e && e[0] == '1') { c->win_stab = 1; c->win_tuned = true; } }
    { const char* e = getenv("GRAAL_LANES"); if (e && e[0] >= '1' && e[0] <= '0' + GRAAL_MAX_LANES) c->n_lanes = e[0] - '0'; }
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_fetch, cudaEventDisableTiming));
    for (int l = 0; l < c->n_lanes; l++) {
        Lane& L = c->lanes[l];
        CUDA_OK(cudaStreamCreateWithFlags(&L.st, cudaStreamNonBlocking));
        CUDA_OK(cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&L.ev_ready, cudaEventDisableTiming));
        for (int i = 0; i < 2; i++) {
            CUDA_OK(cudaStreamCreateWithFlags(&L.side[i], cudaStreamNonBlocking));
            CUDA_OK(cudaEventCreateWithFlags(&L.ev_side[i], cudaEventDisableTiming));
        }
        CUDA_OK(cudaMalloc(&L.ints, 256 * sizeof(int)));
        CUDA_OK(cudaMemset(L.ints, 0, 256 * sizeof(int)));
        CUDA_OK(cudaMalloc(&L.partials, (size_t)48 * c->partial_stride * sizeof(double)));
        CUDA_OK(cudaMalloc(&L.dist_partials, (size_t)GRAAL_N_CANDIDATES * c->partial_stride * sizeof(double)));
    }
    *out = c;
    return 0;
}

static void free_level_scratch(graal_ctx* c) {
    cudaFree(c->cid16_base); cudaFree(c->mid32_base); cudaFree(c->cm_base); c->cm_base = nullptr; cudaFree(c->group_row); cudaFree(c->band_hist); c->band_hist = nullptr; c->cid16_base = nullptr; c->mid32_base = nullptr; c->group_row = nullptr;
    for (int l = 0; l < GRAAL_MAX_LANES; l++) {
        Lane& L = c->lanes[l];
        cudaFree(L.uwin); cudaFree(L.cand_blk); L.uwin = nullptr; L.cand_blk = nullptr;
        cudaFree(L.sub_index); cudaFree(L.chmask); cudaFree(L.chmask2); L.chmask2 = nullptr; cudaFree(L.geo_cand); cudaFree(L.cand_ordrec); cudaFree(L.base_ordrec); cudaFree(L.cand_ordb); cudaFree(L.base_ordb); cudaFree(L.base_ordc); cudaFree(L.rep_in_u);
        L.sub_index = nullptr; L.chmask = nullptr; L.geo_cand = nullptr; L.cand_ordrec = L.base_ordrec = L.cand_ordb = L.base_ordb = L.base_ordc = nullptr; L.rep_in_u = nullptr;
    }
    cudaFree(c->items); cudaFree(c->row_hdr); cudaFree(c->item_hdr); c->item_hdr = nullptr; cudaFree(c->o_start); cudaFree(c->o_sub); cudaFree(c->o_cid); cudaFree(c->o_blk);
    c->items = nullptr; c->row_hdr = nullptr; c->o_start = nullptr; c->o_sub = nullptr; c->o_cid = nullptr; c->o_blk = nullptr; c->n_items = 0;
    cudaFree(c->geo_base); cudaFree(c->order);
    cudaFree(c->cont_len); cudaFree(c->cont_off); cudaFree(c->first_idx); cudaFree(c->map); cudaFree(c->keys);
    cudaFree(c->keys_sorted); cudaFree(c->cub_tmp); cudaFree(c->d_quirky); cudaFree(c->d_accu_idx);
    cudaFree(c->d_dup); cudaFree(c->d_sub_dup); cudaFree(c->d_rep_bins);
    c->d_dup = c->d_sub_dup = nullptr; c->d_rep_bins = nullptr; c->n_rep = 0;
    cudaFree(c->d_tab_lnnorm); cudaFree(c->d_tab_log); cudaFree(c->d_tab_exp); cudaFree(c->d_tab_normd);
    for (int w = 0; w < 2; w++) { cudaFree(c->d_tab_lnf[w]); cudaFree(c->d_tab_f[w]); cudaFree(c->d_tab_lnfu[w]); cudaFree(c->d_tab_lnfu7[w]); c->d_tab_lnfu7[w] = nullptr; cudaFree(c->d_tab_fu[w]); c->d_tab_lnf[w] = nullptr; c->d_tab_f[w] = nullptr; c->d_tab_lnfu[w] = nullptr; c->d_tab_fu[w] = nullptr; }
    c->d_tab_normd = nullptr;
    c->d_tab_lnnorm = nullptr; c->d_tab_log = nullptr; c->d_tab_exp = nullptr;
    cudaFree(c->d_tab_norm); cudaFree(c->d_tab_g[0]); cudaFree(c->d_tab_g[1]); cudaFree(c->d_tab_logg[0]); cudaFree(c->d_tab_logg[1]);
    c->d_accu_idx = nullptr; c->d_tab_norm = nullptr; c->d_tab_g[0] = c->d_tab_g[1] = nullptr; c->d_tab_logg[0] = c->d_tab_logg[1] = nullptr;
    c->geo_base = nullptr; c->order = c->cont_len = c->cont_off = nullptr;
    c->first_idx = c->map = nullptr; c->keys = c->keys_sorted = nullptr; c->cub_tmp = nullptr; c->d_quirky = nullptr;
}

void graal_ctx_destroy(graal_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (int l = 0; l < GRAAL_MAX_LANES; l++) if (c->lanes[l].st) cudaStreamSynchronize(c->lanes[l].st);
    if (c->stream) cudaStreamSynchronize(c->stream);
    free_level_scratch(c);
    for (int l = 0; l < GRAAL_MAX_LANES; l++) {
        Lane& L = c->lanes[l];
        cudaFree(L.ints); cudaFree(L.partials); cudaFree(L.dist_partials);
        if (L.done) cudaEventDestroy(L.done);
        if (L.ev_ready) cudaEventDestroy(L.ev_ready);
        for (int i = 0; i < 2; i++) { if (L.ev_side[i]) cudaEventDestroy(L.ev_side[i]); if (L.side[i]) cudaStreamDestroy(L.side[i]); }
        if (L.st) cudaStreamDestroy(L.st);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_fetch) cudaEventDestroy(c->ev_fetch);
    if (c->h_ncontigs) cudaFreeHost(c->h_ncontigs);
    c->g_prologue.reset();
    for (int i = 0; i < 16; i++) c->graphs[i].reset();
    c->g_stats.reset(); c->g_relabel.reset(); c->g_full.reset(); c->g_full_cached.reset(); c->g_win.reset();
    c->prof.destroy();
    cudaFree(c->d_ints); cudaFree(c->d_stats); cudaFree(c->d_counters); cudaFree(c->d_scalars); cudaFree(c->partials);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int graal_set_stream(graal_ctx* c, void* s) {
    if (!c) return set_err(-1, "null context");
    { int rc = join_lanes(c); if (rc) return rc; }
    c->version++;
    if (c->own_stream && c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    c->stream = (cudaStream_t)s; c->own_stream = false;
    return 0;
}
int graal_sync(graal_ctx* c) {
    if (!c) return set_err(-1, "null context");
    int rc = join_lanes(c); if (rc) return rc;
    CUDA_OK(cudaStreamSynchronize(c->stream));
    // contig ids outside [0, 2 n + 16) seen by the relabel / statistics kernels (slot uploaded from the host, or chains of
    // graal_apply_move without a relabel): reported here, off the step path
    int flag = 0;
    CUDA_OK(cudaMemcpy(&flag, c->d_ints + 2, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) { CUDA_OK(cudaMemset(c->d_ints + 2, 0, sizeof(int))); return set_err(-6, "a contig id outside [0, %d) was met by the relabel / statistics kernels", c->cap); }
    return 0;
}
int graal_join(graal_ctx* c) { if (!c) return set_err(-1, "null context"); return join_lanes(c); }
int64_t graal_launch_count(graal_ctx* c) { return c ? c->launches : -1; }

int graal_profile_enable(graal_ctx* c, int on) {
    if (!c) return set_err(-1, "null context");
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < PROF_KERNELS; k++) { c->prof.resolve(k); }
    c->prof.on = on != 0;
    return 0;
}
int graal_profile_counters(graal_ctx* c, int64_t out[4], int reset) {
    if (!c || !out) return set_err(-1, "null argument");
    CUDA_OK(cudaSetDevice(c->device));
    { int rc = join_lanes(c); if (rc) return rc; }
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaMemcpy(out, c->d_counters, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (reset) CUDA_OK(cudaMemset(c->d_counters, 0, 4 * sizeof(unsigned long long)));
    return 0;
}
int graal_profile_read(graal_ctx* c, int kernel_id, double* total_ms, int64_t* count, int reset) {
    if (!c || kernel_id < 0 || kernel_id >= PROF_KERNELS) return set_err(-1, "bad profile query");
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    c->prof.resolve(kernel_id);
    if (total_ms) *total_ms = c->prof.total_ms[kernel_id];
    if (count) *count = c->prof.count[kernel_id];
    if (reset) { c->prof.total_ms[kernel_id] = 0; c->prof.count[kernel_id] = 0; }
    return 0;
}

int graal_level_bind(graal_ctx* c, int n_frags, int n_new_frags, int n_sub_frags,
                     const int32_t* sub_id, const float* sub_len_kb, const int32_t* sub_accu,
                     const int32_t* collector, const int32_t* dispatcher,
                     const int64_t* rowptr, const void* contacts, int64_t n_contacts, float nfpb) {
    if (!c) return set_err(-1, "null context");
    if (n_frags <= 0 || n_sub_frags <= 0 || n_new_frags < n_frags) return set_err(-1, "bad level sizes");
    if (!sub_id || !sub_len_kb || !sub_accu || !rowptr || (n_contacts > 0 && !contacts)) return set_err(-1, "null level pointer");
    if (n_new_frags != n_frags && (!collector || !dispatcher)) return set_err(-1, "repeat copies need the collector / dispatcher tables");
    CUDA_OK(cudaSetDevice(c->device));
    { int rcj = join_lanes(c); if (rcj) return rcj; }
    CUDA_OK(cudaStreamSynchronize(c->stream));
    c->version++;
    for (int i = 0; i < 16; i++) c->graphs[i].reset();
    c->g_stats.reset(); c->g_relabel.reset(); c->g_full.reset(); c->g_full_cached.reset(); c->g_win.reset();
    c->win_slot = -1; c->contig_bound = -1; c->ncontigs_pending = false; c->fetch_ncontigs = false; c->first_clean = false; c->stats_clean = false; c->g_prologue.reset();
    { const char* e = getenv("GRAAL_WIN_SUB"); const char* e2 = getenv("GRAAL_WIN_STAB"); c->win_tuned = e != nullptr || (e2 && e2[0] == '1'); }
    free_level_scratch(c);
    c->N = n_frags; c->n_new = n_new_frags; c->W = n_sub_frags; c->E = n_contacts; c->nfpb = nfpb;
    c->lv.sub_id = reinterpret_cast<const int4*>(sub_id); c->lv.sub_len = sub_len_kb; c->lv.sub_accu = sub_accu;
    c->collector = collector; c->dispatcher = dispatcher;
    c->rowptr = reinterpret_cast<const long long*>(rowptr); c->contacts = reinterpret_cast<const int2*>(contacts);
    // derived host tables: accu histogram (for G0) and the list of quirky bins (non-uniform accu, Q1)
    std::vector<int> h_sid((size_t)n_frags * 4), h_acc((size_t)n_frags * 3);
    CUDA_OK(cudaMemcpy(h_sid.data(), sub_id, h_sid.size() * sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(h_acc.data(), sub_accu, h_acc.size() * sizeof(int), cudaMemcpyDeviceToHost));
    // duplicated data bins (dispatcher range longer than 1): handled by the repeat path
    std::vector<unsigned char> h_dup((size_t)n_frags, 0); std::vector<int> rep_bins;
    if (n_new_frags != n_frags) {
        std::vector<int> h_disp((size_t)n_frags * 2);
        CUDA_OK(cudaMemcpy(h_disp.data(), dispatcher, h_disp.size() * sizeof(int), cudaMemcpyDeviceToHost));
        for (int b = 0; b < n_frags; b++) if (h_disp[2 * b + 1] - h_disp[2 * b] > 1) { h_dup[b] = 1; rep_bins.push_back(b); }
    }
    {
        std::vector<float> h_len((size_t)n_frags * 3);
        CUDA_OK(cudaMemcpy(h_len.data(), sub_len_kb, h_len.size() * sizeof(float), cudaMemcpyDeviceToHost));
        c->max_bin_kb = 0.0f;
        for (int b = 0; b < n_frags; b++) {
            float l = 0.0f;
            for (int a = 0; a < h_sid[(size_t)b * 4 + 3] && a < 3; a++) l += h_len[(size_t)b * 3 + a];
            c->max_bin_kb = std::max(c->max_bin_kb, l);
        }
    }
    std::map<int, long long> hist; std::vector<int> quirky; long long w_check = 0;
    for (int b = 0; b < n_frags; b++) {
        const int cnt = h_sid[(size_t)b * 4 + 3];
        if (cnt < 1 || cnt > 3) return set_err(-1, "bin %d has %d sub-frags (1..3 supported)", b, cnt);
        if (h_sid[(size_t)b * 4] != (int)w_check) return set_err(-1, "sub-frag ids of bin %d are not consecutive in bin order", b);
        bool q = false;
        for (int a = 0; a < cnt; a++) {
            const int v = h_acc[(size_t)b * 3 + a];
            if (h_sid[(size_t)b * 4 + a] != (int)w_check + a) return set_err(-1, "sub-frag ids of bin %d are not consecutive", b);
            if (v < 0 || v > 46340) return set_err(-1, "accu %d of bin %d outside 0..46340", v, b);
            hist[v] += h_dup[b] ? 0 : 1;
            if (v != h_acc[(size_t)b * 3 + cnt - 1]) q = true;
        }
        if (q && !h_dup[b]) quirky.push_back(b);
        w_check += cnt;
    }
    if (w_check != n_sub_frags) return set_err(-1, "sub-frag count mismatch: %lld vs %d", w_check, n_sub_frags);
    c->accu_hist.assign(hist.begin(), hist.end());
    const int nd = (int)c->accu_hist.size();
    if (nd > MAX_ACCU_VALUES) return set_err(-4, "%d distinct accu values (max %d)", nd, MAX_ACCU_VALUES);
    {
        std::map<int, int> pos; for (int i = 0; i < nd; i++) pos[c->accu_hist[i].first] = i;
        std::vector<unsigned char> idx((size_t)n_frags * 3, 0);
        for (int b = 0; b < n_frags; b++) for (int a = 0; a < h_sid[(size_t)b * 4 + 3]; a++) idx[(size_t)b * 3 + a] = (unsigned char)pos[h_acc[(size_t)b * 3 + a]];
        CUDA_OK(cudaMalloc(&c->d_accu_idx, idx.size()));
        CUDA_OK(cudaMemcpy(c->d_accu_idx, idx.data(), idx.size(), cudaMemcpyHostToDevice));
        std::vector<float> tn((size_t)nd * nd);
        for (int i = 0; i < nd; i++) for (int j = 0; j < nd; j++) tn[(size_t)i * nd + j] = (float)(c->accu_hist[i].first * c->accu_hist[j].first) / nfpb;
        CUDA_OK(cudaMalloc(&c->d_tab_norm, tn.size() * sizeof(float)));
        CUDA_OK(cudaMemcpy(c->d_tab_norm, tn.data(), tn.size() * sizeof(float), cudaMemcpyHostToDevice));
        std::vector<double> ln(tn.size());
        for (size_t i = 0; i < tn.size(); i++) ln[i] = tn[i] > 0.0f ? log((double)tn[i]) : -(double)INFINITY;
        CUDA_OK(cudaMalloc(&c->d_tab_lnnorm, ln.size() * sizeof(double)));
        CUDA_OK(cudaMemcpy(c->d_tab_lnnorm, ln.data(), ln.size() * sizeof(double), cudaMemcpyHostToDevice));
        std::vector<double> tl(256), te(32);
        for (int i = 0; i < 128; i++) { const double cc = 1.0 + (i + 0.5) / 128.0; tl[2 * i] = 1.0 / cc; tl[2 * i + 1] = log(cc); }
        for (int j = 0; j < 32; j++) te[j] = exp2((double)j / 32.0);
        CUDA_OK(cudaMalloc(&c->d_tab_log, 128 * sizeof(double2)));
        CUDA_OK(cudaMemcpy(c->d_tab_log, tl.data(), 256 * sizeof(double), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMalloc(&c->d_tab_exp, 32 * sizeof(double)));
        CUDA_OK(cudaMemcpy(c->d_tab_exp, te.data(), 32 * sizeof(double), cudaMemcpyHostToDevice));
        std::vector<double> tnd(tn.begin(), tn.end());
        CUDA_OK(cudaMalloc(&c->d_tab_normd, tnd.size() * sizeof(double)));
        CUDA_OK(cudaMemcpy(c->d_tab_normd, tnd.data(), tnd.size() * sizeof(double), cudaMemcpyHostToDevice));
        for (int w = 0; w < 2; w++) {
            CUDA_OK(cudaMalloc(&c->d_tab_lnf[w], (size_t)(LAW_NODES + 2) * sizeof(int4)));
            CUDA_OK(cudaMalloc(&c->d_tab_lnfu[w], (size_t)(LAW_NODES + 2) * sizeof(int4)));
            CUDA_OK(cudaMalloc(&c->d_tab_lnfu7[w], (size_t)LAW7_NODES * sizeof(int4)));
            CUDA_OK(cudaMalloc(&c->d_tab_fu[w], (size_t)(LAW_NODES + 2) * sizeof(int4)));
            CUDA_OK(cudaMalloc(&c->d_tab_f[w], (size_t)(LAW_NODES + 2) * sizeof(int4)));
        }
        for (int w = 0; w < 2; w++) {
            CUDA_OK(cudaMalloc(&c->d_tab_g[w], tn.size() * sizeof(float)));
            CUDA_OK(cudaMalloc(&c->d_tab_logg[w], tn.size() * sizeof(double)));
        }
        c->lv.accu_idx = c->d_accu_idx;
    }
    c->have_params = false;
    c->n_rep = (int)rep_bins.size();
    c->lv.dup = nullptr; c->lv.n_data = n_frags;
    if (c->n_rep > 0) {
        std::vector<unsigned char> sub_dup((size_t)n_sub_frags, 0);
        for (int b = 0; b < n_frags; b++) if (h_dup[b]) for (int a = 0; a < h_sid[(size_t)b * 4 + 3]; a++) sub_dup[h_sid[(size_t)b * 4] + a] = 1;
        CUDA_OK(cudaMalloc(&c->d_dup, h_dup.size())); CUDA_OK(cudaMemcpy(c->d_dup, h_dup.data(), h_dup.size(), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMalloc(&c->d_sub_dup, sub_dup.size())); CUDA_OK(cudaMemcpy(c->d_sub_dup, sub_dup.data(), sub_dup.size(), cudaMemcpyHostToDevice));
        for (int l = 0; l < c->n_lanes; l++) { CUDA_OK(cudaMalloc(&c->lanes[l].rep_in_u, h_dup.size())); CUDA_OK(cudaMemset(c->lanes[l].rep_in_u, 0, h_dup.size())); }
        CUDA_OK(cudaMalloc(&c->d_rep_bins, rep_bins.size() * sizeof(int)));
        CUDA_OK(cudaMemcpy(c->d_rep_bins, rep_bins.data(), rep_bins.size() * sizeof(int), cudaMemcpyHostToDevice));
        c->lv.dup = c->d_dup;
    }
    c->n_quirky = (int)quirky.size();
    if (c->n_quirky) {
        CUDA_OK(cudaMalloc(&c->d_quirky, quirky.size() * sizeof(int)));
        CUDA_OK(cudaMemcpy(c->d_quirky, quirky.data(), quirky.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    const int n = n_new_frags;
    c->cap = 2 * n + 16;
    CUDA_OK(cudaMalloc(&c->geo_base, (size_t)c->W * sizeof(Geo)));
    CUDA_OK(cudaMalloc(&c->cid16_base, ((size_t)c->W + 16) * sizeof(unsigned short)));
    CUDA_OK(cudaMemset(c->cid16_base, 0, ((size_t)c->W + 16) * sizeof(unsigned short)));
    CUDA_OK(cudaMalloc(&c->cm_base, (size_t)c->W * sizeof(int2)));
    CUDA_OK(cudaMemset(c->cm_base, 0, (size_t)c->W * sizeof(int2)));
    CUDA_OK(cudaMalloc(&c->mid32_base, (size_t)c->W * sizeof(float)));
    for (int l = 0; l < c->n_lanes; l++) {
        Lane& L = c->lanes[l];
        CUDA_OK(cudaMalloc(&L.chmask, (size_t)c->W * sizeof(unsigned)));
        CUDA_OK(cudaMalloc(&L.chmask2, (size_t)c->W * sizeof(unsigned)));
        CUDA_OK(cudaMemset(L.chmask2, 0, (size_t)c->W * sizeof(unsigned)));
        CUDA_OK(cudaMalloc(&L.cand_ordrec, (size_t)N_BAND_ROWS * n * sizeof(int4)));
        CUDA_OK(cudaMalloc(&L.base_ordrec, (size_t)n * sizeof(int4)));
        CUDA_OK(cudaMemset(L.cand_ordrec, 0, (size_t)N_BAND_ROWS * n * sizeof(int4)));
        CUDA_OK(cudaMemset(L.base_ordrec, 0, (size_t)n * sizeof(int4)));
        CUDA_OK(cudaMalloc(&L.cand_ordb, (size_t)N_BAND_ROWS * n * sizeof(int4)));
        CUDA_OK(cudaMalloc(&L.base_ordb, (size_t)n * sizeof(int4)));
        CUDA_OK(cudaMalloc(&L.base_ordc, (size_t)n * sizeof(int4)));
        CUDA_OK(cudaMemset(L.cand_ordb, 0, (size_t)N_BAND_ROWS * n * sizeof(int4)));
        CUDA_OK(cudaMemset(L.base_ordb, 0, (size_t)n * sizeof(int4)));
        CUDA_OK(cudaMemset(L.base_ordc, 0, (size_t)n * sizeof(int4)));
        CUDA_OK(cudaMalloc(&L.geo_cand, (size_t)GRAAL_N_CANDIDATES * c->W * sizeof(Geo)));
        CUDA_OK(cudaMemset(L.geo_cand, 0, (size_t)GRAAL_N_CANDIDATES * c->W * sizeof(Geo)));
        CUDA_OK(cudaMalloc(&L.uwin, (size_t)c->W * sizeof(int2)));
        CUDA_OK(cudaMemset(L.uwin, 0, (size_t)c->W * sizeof(int2)));
        CUDA_OK(cudaMalloc(&L.cand_blk, (size_t)GRAAL_N_CANDIDATES * (n / 32 + 1) * sizeof(int2)));
        CUDA_OK(cudaMalloc(&L.sub_index, (size_t)n * sizeof(int)));
        CUDA_OK(cudaMemset(L.sub_index, 0, (size_t)n * sizeof(int)));
        L.pending = false; L.cand_first = -1;
    }
    CUDA_OK(cudaMalloc(&c->band_hist, 16 * GRAAL_N_CANDIDATES * sizeof(double)));
    CUDA_OK(cudaMemset(c->band_hist, 0, 16 * GRAAL_N_CANDIDATES * sizeof(double)));
    c->band_slot = -1;
    CUDA_OK(cudaMalloc(&c->order, (size_t)n * sizeof(int)));
    CUDA_OK(cudaMalloc(&c->cont_len, (size_t)c->cap * sizeof(int)));
    CUDA_OK(cudaMalloc(&c->cont_off, (size_t)c->cap * sizeof(int)));
    CUDA_OK(cudaMalloc(&c->first_idx, (size_t)c->cap * sizeof(int)));
    CUDA_OK(cudaMalloc(&c->map, (size_t)c->cap * sizeof(int)));
    CUDA_OK(cudaMalloc(&c->keys, (size_t)c->cap * sizeof(unsigned long long)));
    CUDA_OK(cudaMalloc(&c->keys_sorted, (size_t)c->cap * sizeof(unsigned long long)));
    size_t b1 = 0, b2 = 0;
    c->key_bits = 1; while ((1ll << c->key_bits) < (long long)n + 2) c->key_bits++;      // lengths 0..n and the sentinel
    cub::DeviceRadixSort::SortPairs(nullptr, b1, reinterpret_cast<unsigned*>(c->keys), reinterpret_cast<unsigned*>(c->keys_sorted),
                                    reinterpret_cast<unsigned*>(c->keys), reinterpret_cast<unsigned*>(c->keys_sorted), c->cap, 0, c->key_bits);
    cub::DeviceScan::ExclusiveSum(nullptr, b2, c->cont_len, c->cont_off, c->cap);
    c->cub_tmp_bytes = std::max(b1, b2);
    CUDA_OK(cudaMalloc(&c->cub_tmp, c->cub_tmp_bytes));
    c->geo_base_slot = -1; c->first_idx_slot = -1;
    c->n_groups = (int)((c->E + GROUP - 1) / GROUP);
    if (c->n_groups > 0) {
        CUDA_OK(cudaMalloc(&c->group_row, (size_t)c->n_groups * sizeof(int)));
        k_group_rows<<<(c->n_groups + 255) / 256, 256, 0, c->stream>>>(c->rowptr, c->E, c->W, c->n_groups, c->group_row); CHECK_LAUNCH(c);
    }
    // work items of the windowed contact pass: rows cut into chunks of <= ITEM_CHUNK entries
    CUDA_OK(cudaMalloc(&c->row_hdr, (size_t)c->W * sizeof(int4)));
    CUDA_OK(cudaMalloc(&c->o_start, ((size_t)n + 1) * sizeof(int)));
    CUDA_OK(cudaMalloc(&c->o_sub, (size_t)n * sizeof(int2)));
    CUDA_OK(cudaMalloc(&c->o_cid, (size_t)n * sizeof(int)));
    CUDA_OK(cudaMalloc(&c->o_blk, ((size_t)n / 32 + 1) * sizeof(int2)));
    if (c->E > 0) {
        int* cnt = nullptr; int* off = nullptr; void* tmp = nullptr; size_t tb = 0;
        CUDA_OK(cudaMalloc(&cnt, (size_t)c->W * sizeof(int)));
        CUDA_OK(cudaMalloc(&off, (size_t)c->W * sizeof(int)));
        k_item_count<<<nblk(c->W, 256), 256, 0, c->stream>>>(c->rowptr, c->W, cnt); CHECK_LAUNCH(c);
        cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, off, c->W, c->stream);
        CUDA_OK(cudaMalloc(&tmp, tb));
        CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, off, c->W, c->stream));
        int last_off = 0, last_cnt = 0;
        CUDA_OK(cudaMemcpyAsync(&last_off, off + c->W - 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaMemcpyAsync(&last_cnt, cnt + c->W - 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        c->n_items = last_off + last_cnt;
        if (c->n_items > 0) {
            CUDA_OK(cudaMalloc(&c->items, (size_t)c->n_items * sizeof(int4)));
            CUDA_OK(cudaMalloc(&c->item_hdr, (size_t)c->n_items * sizeof(int4)));
            k_item_fill<<<nblk(c->W, 256), 256, 0, c->stream>>>(c->rowptr, c->W, off, c->items); CHECK_LAUNCH(c);
        }
        CUDA_OK(cudaStreamSynchronize(c->stream));
        cudaFree(cnt); cudaFree(off); cudaFree(tmp);
    }
    // level constant: sum of lf(ob)
    c->lf_total = 0.0;
    if (c->E > 0) {
        const int grid = c->partial_stride;
        k_lf_total<<<grid, 256, 0, c->stream>>>(c->contacts, c->E, c->rowptr, c->W, c->d_sub_dup, c->partials);
        CHECK_LAUNCH(c);
        k_reduce_partials<<<1, 256, 0, c->stream>>>(c->partials, grid, 0, 1.0, c->d_scalars + 15, 0);
        CHECK_LAUNCH(c);
        CUDA_OK(cudaMemcpyAsync(&c->lf_total, c->d_scalars + 15, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        k_ob_total<<<grid, 256, 0, c->stream>>>(c->contacts, c->E, c->rowptr, c->W, c->d_sub_dup, c->partials);
        CHECK_LAUNCH(c);
        k_reduce_partials<<<1, 256, 0, c->stream>>>(c->partials, grid, 0, 1.0, c->d_scalars + 14, 0);
        CHECK_LAUNCH(c);
        CUDA_OK(cudaMemcpyAsync(&c->ob_total, c->d_scalars + 14, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int graal_set_params(graal_ctx* c, const float q[8]) {
    if (!c || !q) return set_err(-1, "null argument");
    c->p.kuhn = q[0]; c->p.lm = q[1]; c->p.c1 = q[2]; c->p.slope = q[3]; c->p.d = q[4];
    c->p.d_max = q[5]; c->p.fact = q[6]; c->p.v_inter = q[7]; c->p.nfpb = c->nfpb;
    if (!c->d_tab_norm) return set_err(-1, "bind the level before setting parameters");
    CUDA_OK(cudaSetDevice(c->device));
    int rc = join_lanes(c); if (rc) return rc;          // pending proposals still read the tables being replaced
    rc = upload_tables(c, c->p, 0); if (rc) return rc;
    c->have_params = true; c->band_slot = -1; c->version++;
    return 0;
}

int graal_set_math_mode(graal_ctx* c, int mode) {
    if (!c || mode < 0 || mode > 2) return set_err(-1, "math mode must be 0 (float32 chain), 1 (log-space float64) or 2 (tabulated law)");
    c->math_mode = mode; c->p.mode = mode; c->band_slot = -1; c->version++;
    return 0;
}

int graal_state_bind(graal_ctx* c, int32_t* base, int ld, int n_slots) {
    if (!c || !base) return set_err(-1, "null argument");
    if (c->n_new <= 0) return set_err(-1, "bind the level first");
    if (ld < c->n_new || n_slots < 1) return set_err(-1, "bad slot geometry (ld %d < n %d)", ld, c->n_new);
    { int rcj = join_lanes(c); if (rcj) return rcj; }
    c->version++;
    c->slots = base; c->ld = ld; c->n_slots = n_slots; c->geo_base_slot = -1; c->band_slot = -1; c->first_idx_slot = -1;
    c->contig_bound = -1; c->ncontigs_pending = false; c->fetch_ncontigs = false;
    return 0;
}

#define NEED_STATE(c) do { if (!(c) || !(c)->slots) return set_err(-1, "state not bound"); CUDA_OK(cudaSetDevice((c)->device)); \
                           { int rcj_ = join_lanes(c); if (rcj_) return rcj_; } } while (0)
#define NEED_SLOT(c, s) do { if ((s) < 0 || (s) >= (c)->n_slots) return set_err(-1, "slot %d out of range", (s)); } while (0)

int graal_relabel_contigs(graal_ctx* c, int slot, int32_t* d_max_id) {
    NEED_STATE(c); NEED_SLOT(c, slot);
    const int n = c->n_new, cap = c->cap, ld = c->ld;
    int* s = slot_ptr(c, slot);
    cudaStream_t st = c->stream;
    const bool have_first = c->first_idx_slot == slot;      // graal_state_stats of the same state leaves the table behind
    auto enqueue = [&]() -> int {
        c->prof.begin(GRAAL_K_RELABEL, st);
        if (!have_first) {
            k_fill_int<<<nblk(cap, 256), 256, 0, st>>>(c->first_idx, cap, INT_MAX); CHECK_LAUNCH(c);
            k_first_index<<<nblk(n, 256), 256, 0, st>>>(s + F_ID_C * ld, n, cap, c->first_idx, c->d_ints + 2); CHECK_LAUNCH(c);
        }
        unsigned* k_in = reinterpret_cast<unsigned*>(c->keys); unsigned* v_in = k_in + cap;
        unsigned* k_out = reinterpret_cast<unsigned*>(c->keys_sorted); unsigned* v_out = k_out + cap;
        const unsigned sentinel = (1u << c->key_bits) - 1u;
        k_relabel_keys<<<nblk(cap, 256), 256, 0, st>>>(c->first_idx, s + F_L_CONT * ld, cap, sentinel, k_in, v_in); CHECK_LAUNCH(c);
        size_t tb = c->cub_tmp_bytes;
        CUDA_OK(cub::DeviceRadixSort::SortPairs(c->cub_tmp, tb, k_in, k_out, v_in, v_out, cap, 0, c->key_bits, st)); c->launches += 2 + (c->key_bits + 7) / 8;
        k_relabel_map<<<nblk(cap, 256), 256, 0, st>>>(k_out, v_out, cap, sentinel, c->map, c->d_ints + 1); CHECK_LAUNCH(c);
        k_relabel_apply<<<nblk(n, 256), 256, 0, st>>>(s + F_ID_C * ld, n, cap, c->map, c->d_ints + 1, c->d_ints + 0, d_max_id); CHECK_LAUNCH(c);
        c->prof.end(GRAAL_K_RELABEL, st);
        return 0;
    };
    { const int rc = run_graphed(c, c->g_relabel, slot, (long long)(uintptr_t)d_max_id, have_first ? 1 : 0, enqueue); if (rc) return rc; }
    if (c->geo_base_slot == slot) c->geo_base_slot = -1;
    c->first_idx_slot = -1;                  // indexed by the OLD contig ids
    c->first_clean = false;
    if (c->bound_slot != slot) { c->bound_slot = slot; c->contig_bound = -1; c->fetch_ncontigs = false; }
    c->bound_commits = 0; c->ncontigs_pending = true;      // d_ints[1] = n_contigs of this slot: read back by the next graal_fetch
    return 0;
}

// graal_state_stats + graal_relabel_contigs of one slot in three launches when the host knows a bound on the contig ids
// (see k_prologue_scan); the general sequences otherwise.
int graal_stats_relabel(graal_ctx* c, int slot, double* d_stats_out, int32_t* d_max_id) {
    NEED_STATE(c); NEED_SLOT(c, slot);
    if (!d_stats_out) return set_err(-1, "null output");
    const bool fused = c->fused_prologue && !c->prof.on && c->bound_slot == slot && c->contig_bound >= 0 && c->contig_bound <= RL_MAX;
    if (!fused) {
        int rc = graal_state_stats(c, slot, d_stats_out); if (rc) return rc;
        rc = graal_relabel_contigs(c, slot, d_max_id); if (rc) return rc;
        if (c->fused_prologue && !c->prof.on && c->contig_bound < 0 && c->h_ncontigs) {
            // bound unknown (first step after an upload / a raw mutation): ONE blocking read of the contig count seeds it, the
            // following steps take the fused path without any (a run that never fetches would otherwise never learn it)
            CUDA_OK(cudaMemcpyAsync(c->h_ncontigs, c->d_ints + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            CUDA_OK(cudaStreamSynchronize(c->stream));
            c->contig_bound = *c->h_ncontigs; c->bound_commits = 0; c->ncontigs_pending = false;
        }
        return 0;
    }
    const int n = c->n_new, ld = c->ld, cap = c->cap;
    int* s = slot_ptr(c, slot);
    cudaStream_t st = c->stream;
    const bool fill = !c->first_clean, init = !c->stats_clean;
    auto enqueue = [&]() -> int {
        if (fill) { k_fill_int<<<nblk(cap, 256), 256, 0, st>>>(c->first_idx, cap, INT_MAX); CHECK_LAUNCH(c); }
        if (init) { k_init_stats<<<1, 1, 0, st>>>(c->d_stats); CHECK_LAUNCH(c); }
        k_prologue_scan<<<std::min(nblk(n, 256), c->n_sm * 4), 256, 0, st>>>(s, ld, n, c->first_idx, c->d_stats, c->d_ints + 2); CHECK_LAUNCH(c);
        k_prologue_rank<<<1, 1024, 0, st>>>(c->first_idx, s + F_L_CONT * ld, std::min(cap, RL_RANGE), c->map, c->d_ints, c->d_stats, d_stats_out, d_max_id); CHECK_LAUNCH(c);
        k_relabel_apply<<<nblk(n, 256), 256, 0, st>>>(s + F_ID_C * ld, n, cap, c->map, c->d_ints + 1, c->d_ints + 0, d_max_id); CHECK_LAUNCH(c);
        return 0;
    };
    { const int rc = run_graphed(c, c->g_prologue, slot, (long long)(uintptr_t)d_stats_out ^ ((long long)(uintptr_t)d_max_id << 1), (fill ? 1 : 0) | (init ? 2 : 0), enqueue); if (rc) return rc; }
    c->first_clean = true; c->stats_clean = true;
    if (c->geo_base_slot == slot) c->geo_base_slot = -1;
    c->first_idx_slot = -1;
    c->bound_commits = 0; c->ncontigs_pending = true;
    return 0;
}

int graal_apply_move(graal_ctx* c, int src_slot, int dst_slot, int op, int id_fA, int id_fB, int aux,
                     int max_id_in, int32_t* d_max_id_out) {
    NEED_STATE(c); NEED_SLOT(c, src_slot); NEED_SLOT(c, dst_slot);
    if (op < 0 || op >= GRAAL_N_OPS) return set_err(-1, "unknown op %d", op);
    const int n = c->n_new;
    if (id_fA < 0 || id_fA >= n || id_fB < 0 || id_fB >= n) return set_err(-1, "bin id out of range");
    if (src_slot == dst_slot) return set_err(-1, "in-place moves are not supported (src == dst)");
    k_apply_move<<<nblk(n, 128), 128, 0, c->stream>>>(slot_ptr(c, src_slot), slot_ptr(c, dst_slot), c->ld, n, op, id_fA, id_fB, aux, max_id_in);
    CHECK_LAUNCH(c);
    if (d_max_id_out) {
        k_set_int<<<1, 1, 0, c->stream>>>(d_max_id_out, INT_MIN); CHECK_LAUNCH(c);
        k_max_field<<<std::min(nblk(n, 256), c->n_sm * 4), 256, 0, c->stream>>>(slot_ptr(c, dst_slot) + F_ID_C * c->ld, n, d_max_id_out);
        CHECK_LAUNCH(c);
    }
    if (c->geo_base_slot == dst_slot) c->geo_base_slot = -1;
    if (c->first_idx_slot == dst_slot) c->first_idx_slot = -1;
    if (c->band_slot == dst_slot) c->band_slot = -1;
    if (c->bound_slot == dst_slot) { c->contig_bound = -1; c->ncontigs_pending = false; c->fetch_ncontigs = false; }
    return 0;
}

int graal_build_candidates(graal_ctx* c, int src_slot, int first_dst_slot, int id_fA, int id_fB, int max_id, unsigned mask) {
    NEED_STATE(c); NEED_SLOT(c, src_slot); NEED_SLOT(c, first_dst_slot); NEED_SLOT(c, first_dst_slot + GRAAL_N_CANDIDATES - 1);
    const int n = c->n_new;
    if (id_fA < 0 || id_fA >= n || id_fB < 0 || id_fB >= n) return set_err(-1, "bin id out of range");
    if (src_slot >= first_dst_slot && src_slot < first_dst_slot + GRAAL_N_CANDIDATES) return set_err(-1, "source slot inside the destination range");
    c->prof.begin(GRAAL_K_BUILD, c->stream);
    k_build_candidates<<<nblk(n, 128), 128, 0, c->stream>>>(slot_ptr(c, src_slot), slot_ptr(c, first_dst_slot), slot_stride(c), c->ld, n,
                                                           id_fA, id_fB, c->d_ints + 0, max_id, mask & 0x1FFFu);
    CHECK_LAUNCH(c);
    c->prof.end(GRAAL_K_BUILD, c->stream);
    if (c->geo_base_slot >= first_dst_slot && c->geo_base_slot < first_dst_slot + GRAAL_N_CANDIDATES) c->geo_base_slot = -1;
    if (c->first_idx_slot >= first_dst_slot && c->first_idx_slot < first_dst_slot + GRAAL_N_CANDIDATES) c->first_idx_slot = -1;
    if (c->band_slot >= first_dst_slot && c->band_slot < first_dst_slot + GRAAL_N_CANDIDATES) c->band_slot = -1;
    if (c->bound_slot >= first_dst_slot && c->bound_slot < first_dst_slot + GRAAL_N_CANDIDATES) { c->contig_bound = -1; c->ncontigs_pending = false; c->fetch_ncontigs = false; }
    return 0;
}

int graal_commit(graal_ctx* c, int dst_slot, int src_slot) {
    NEED_STATE(c); NEED_SLOT(c, src_slot); NEED_SLOT(c, dst_slot);
    if (src_slot == dst_slot) return 0;
    k_apply_move<<<nblk(c->n_new, 128), 128, 0, c->stream>>>(slot_ptr(c, src_slot), slot_ptr(c, dst_slot), c->ld, c->n_new, GRAAL_OP_COPY, 0, 0, 0, 0);
    CHECK_LAUNCH(c);
    if (c->geo_base_slot == dst_slot) c->geo_base_slot = -1;
    if (c->first_idx_slot == dst_slot) c->first_idx_slot = -1;
    if (c->band_slot == dst_slot) c->band_slot = -1;
    // a committed candidate of a relabelled genome adds at most 3 contig ids (eject + split insert / two splits)
    if (c->bound_slot == dst_slot) { c->bound_commits++; c->commits_total++; if (c->contig_bound >= 0) c->contig_bound += 3; }
    return 0;
}

static int ensure_base_geometry(graal_ctx* c, int slot) {
    if (c->geo_base_slot == slot) return 0;
    k_geometry_all<<<nblk(c->n_new, 128), 128, 0, c->stream>>>(slot_ptr(c, slot), c->ld, c->n_new, c->lv, c->geo_base, c->cid16_base, c->mid32_base, c->cm_base);
    CHECK_LAUNCH(c);
    c->geo_base_slot = slot; c->geo_epoch++;
    return 0;
}

// Position order of `slot` (c->order, cont_off / cont_len) and the row records of the windowed passes (row_hdr, item_hdr)
// for the band d_max; the geometry of the slot must be current.  Cached per (slot, geometry, d_max).
static bool windows_current(const graal_ctx* c, int slot, float d_max) {
    return c->win_slot == slot && c->geo_base_slot == slot && c->win_epoch == c->geo_epoch && c->win_dmax == d_max;
}
static int enqueue_base_windows(graal_ctx* c, int slot, float d_max) {
    cudaStream_t st = c->stream;
    const int n = c->n_new, ld = c->ld;
    int* s = slot_ptr(c, slot);
    int* bad = c->d_ints + 4;
    c->prof.begin(GRAAL_K_FULL_WINDOWS, st);
    CUDA_OK(cudaMemsetAsync(c->cont_len, 0, (size_t)c->cap * sizeof(int), st));
    k_contig_lengths<<<nblk(n, 256), 256, 0, st>>>(s, ld, n, c->cap, c->cont_len); CHECK_LAUNCH(c);
    size_t tb = c->cub_tmp_bytes;
    CUDA_OK(cub::DeviceScan::ExclusiveSum(c->cub_tmp, tb, c->cont_len, c->cont_off, c->cap, st)); c->launches += 2;
    k_order_init<<<nblk(n, 256), 256, 0, st>>>(n, c->W, c->order, c->o_start, c->o_sub, c->o_cid, bad); CHECK_LAUNCH(c);
    k_order_fill2<<<nblk(n, 256), 256, 0, st>>>(s, ld, n, c->cap, c->cont_off, c->lv, c->order, c->o_start, c->o_sub, c->o_cid, bad); CHECK_LAUNCH(c);
    k_hull_blocks<<<nblk(n, 256), 256, 0, st>>>(c->o_sub, n, c->o_blk); CHECK_LAUNCH(c);
    const double dm = (double)d_max * 1000.0 * 1.0001 + 100.0;
    const long long dmax_bp = dm < 4.0e18 ? (long long)ceil(dm) : (long long)4.0e18;
    k_windows<<<nblk(n, 256), 256, 0, st>>>(c->order, c->o_start, c->o_sub, c->o_cid, c->o_blk, c->cont_off, c->cont_len, s, ld, n, c->W,
                                           dmax_bp, c->cm_base, c->row_hdr, bad); CHECK_LAUNCH(c);
    if (c->n_items > 0) { k_item_hdr<<<nblk(c->n_items, 256), 256, 0, st>>>(c->items, c->n_items, c->row_hdr, c->item_hdr); CHECK_LAUNCH(c); }
    c->prof.end(GRAAL_K_FULL_WINDOWS, st);
    return 0;
}
static int ensure_base_windows(graal_ctx* c, int slot, float d_max) {
    if (windows_current(c, slot, d_max)) return 0;
    union { float f; unsigned u; } dm; dm.f = d_max;
    const int rc = run_graphed(c, c->g_win, slot, (long long)dm.u, 0, [&]() -> int { return enqueue_base_windows(c, slot, d_max); });
    if (rc) return rc;
    c->win_slot = slot; c->win_epoch = c->geo_epoch; c->win_dmax = d_max;
    return 0;
}

int graal_full_loglik(graal_ctx* c, int slot, const float* p_override, double* d_out) {
    if (!c || !c->slots) return set_err(-1, "state not bound");
    CUDA_OK(cudaSetDevice(c->device));
    NEED_SLOT(c, slot);
    {   // The full likelihood only reads the slot and the base geometry: it may run beside pending proposals
        // when they were scored against this very slot (geometry current) and it is not one of their candidates.
        bool join = c->geo_base_slot != slot || p_override != nullptr || c->prof.on;
        for (int l = 0; l < c->n_lanes; l++)
            if (c->lanes[l].pending && slot >= c->lanes[l].cand_first && slot < c->lanes[l].cand_first + GRAAL_N_CANDIDATES) join = true;
        if (join) { int rcj = join_lanes(c); if (rcj) return rcj; }
    }
    if (!d_out) return set_err(-1, "null output");
    if (!c->have_params && !p_override) return set_err(-1, "parameters not set");
    Params p = c->p;
    if (p_override) { p.kuhn = p_override[0]; p.lm = p_override[1]; p.c1 = p_override[2]; p.slope = p_override[3];
                      p.d = p_override[4]; p.d_max = p_override[5]; p.fact = p_override[6]; p.v_inter = p_override[7]; p.nfpb = c->nfpb;
                      int rc0 = upload_tables(c, p, 1); if (rc0) return rc0; }
    cudaStream_t st = c->stream;
    const int n = c->n_new, ld = c->ld;
    int* s = slot_ptr(c, slot);
    int rc = ensure_base_geometry(c, slot); if (rc) return rc;
    const int ps = c->partial_stride;
    // contacts
    const double g0 = host_g0(c, p);
    // uniform-accu levels with a finite log g: the trans / clamped entries are a level constant
    const int nd = (int)c->accu_hist.size();
    double lg_uniform = 0.0;
    bool uniform = false;
    if (nd == 1 && c->n_rep == 0) {
        const float gg = g_clamp(c->accu_hist[0].first * c->accu_hist[0].first, p.v_inter, p.nfpb);
        if (gg > 0.0f) { uniform = true; lg_uniform = log((double)gg); }
    }
    // shared-memory classification needs W * 2 B + the queues in one CTA per SM
    const size_t fcs_smem = (size_t)(FCS_THREADS / 32) * QCAP * sizeof(Pending) + (((size_t)c->W + 7) / 8) * 16;
    const bool use_smem = uniform && c->smem_cid && fcs_smem + 1024 <= (size_t)c->smem_optin;
    const bool use_win = uniform && p.mode == 2 && p.fu_ok && c->full_win && !use_smem && c->n_items > 0;
    int g1 = 0;
    typedef void (*win_fn)(const int4*, const int4*, int, const int2*, const int2*, const int*, int, const Geo*, const FastLaw, const Params, double, double*);
    win_fn k_win = k_full_contacts_win<8, 4, 4>;
    switch (c->win_unroll * 100 + c->win_minb * 10 + c->win_sub) {
        case 441: k_win = k_full_contacts_win<4, 4, 1>; break;  case 442: k_win = k_full_contacts_win<4, 4, 2>; break;
        case 444: k_win = k_full_contacts_win<4, 4, 4>; break;  case 454: k_win = k_full_contacts_win<4, 5, 4>; break;
        case 841: k_win = k_full_contacts_win<8, 4, 1>; break;  case 842: k_win = k_full_contacts_win<8, 4, 2>; break;
        case 844: k_win = k_full_contacts_win<8, 4, 4>; break;  case 848: k_win = k_full_contacts_win<8, 4, 8>; break;
        case 832: k_win = k_full_contacts_win<8, 3, 2>; break;  case 834: k_win = k_full_contacts_win<8, 3, 4>; break;
        case 838: k_win = k_full_contacts_win<8, 3, 8>; break;  case 852: k_win = k_full_contacts_win<8, 5, 2>; break;
        case 854: k_win = k_full_contacts_win<8, 5, 4>; break;
        default: break;
    }
    if (c->E > 0 && use_win) {
        int fc_blocks = 0;
        CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fc_blocks, k_win, 256, 0));
        g1 = (int)std::min<long long>(std::min(ps, c->n_sm * std::max(1, fc_blocks)), ((long long)c->n_items + 7) / 8);   // one resident wave
    } else if (c->E > 0) {
        if (use_smem) {
            CUDA_OK(cudaFuncSetAttribute(k_full_contacts_uniform<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fcs_smem));
            g1 = std::min(ps, c->n_sm);
        } else {
            int fc_blocks = 0;
            if (uniform && p.mode == 2) { CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fc_blocks, k_full_contacts_direct, 256, 0)); }
            else if (uniform) { CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fc_blocks, k_full_contacts_uniform<false>, 256, 0)); }
            else { CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fc_blocks, k_full_contacts, 256, 0)); }
            g1 = (int)std::min<long long>(std::min(ps, c->n_sm * std::max(1, fc_blocks)), (c->E + 1023) / 1024);   // one resident wave
        }
    }
    if (c->n_quirky > ps) return set_err(-5, "too many quirky bins (%d > %d)", c->n_quirky, ps);
    if (use_win && g1 > 0) { rc = ensure_base_windows(c, slot, p.d_max); if (rc) return rc; }
    auto win_variant = [&](int sub, int stab) -> win_fn {
        if (stab) return (sub == 2) ? k_full_contacts_win<8, 4, 2, LAW7_M> : k_full_contacts_win<8, 4, 4, LAW7_M>;
        return (sub == 2) ? k_full_contacts_win<8, 4, 2> : k_full_contacts_win<8, 4, 4>;
    };
    auto win_law = [&](int stab) {
        FastLaw fl; fl.zlo = p.fu_zlo; fl.zspan = p.fu_zspan;
        if (stab) { fl.tab = p.t_lnfu7; fl.smin_bits = p.fu7_smin; fl.span = p.fu7_span; }
        else { fl.tab = p.t_lnfu; fl.smin_bits = LAW_SMIN_BITS; fl.span = p.fu_span; }
        return fl;
    };
    const bool stab_ok = c->win_stab_allowed && p.fu7_span > 0u;
    if (use_win && g1 > 0 && !c->win_tuned && !p_override && c->win_unroll == 8 && c->win_minb == 4) {
        // near-diagonal lists want 4 loads per exact-path group, lists dominated by far entries 2; the law table from shared
        // memory or through L1: the four combinations are timed once on this level
        cudaEvent_t e0, e1; CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
        float best = 1e30f; int best_sub = c->win_sub, best_stab = 0;
        for (int stab = 0; stab <= (stab_ok ? 1 : 0); stab++)
        for (int sub = 2; sub <= 4; sub += 2) {
            win_fn kf = win_variant(sub, stab);
            const FastLaw fl = win_law(stab);
            float ms = 1e30f;
            for (int rep = 0; rep < 3; rep++) {
                CUDA_OK(cudaEventRecord(e0, st));
                kf<<<g1, 256, 0, st>>>(c->items, c->item_hdr, c->n_items, c->contacts, c->cm_base, c->d_ints + 4, c->W, c->geo_base, fl, p, lg_uniform, c->partials);
                CHECK_LAUNCH(c);
                CUDA_OK(cudaEventRecord(e1, st));
                CUDA_OK(cudaEventSynchronize(e1));
                float t = 0.f; CUDA_OK(cudaEventElapsedTime(&t, e0, e1));
                if (rep > 0) ms = std::min(ms, t);
            }
            if (ms < (stab ? 0.97f * best : best)) { best = ms; best_sub = sub; best_stab = stab; }     // the coarse table only for a real gain (> 3 %)
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        c->win_sub = best_sub; c->win_stab = best_stab; c->win_tuned = true;
        c->g_full.reset(); c->g_full_cached.reset();
    }
    int use_stab = (c->win_stab && stab_ok && c->win_unroll == 8 && c->win_minb == 4 && (c->win_sub == 2 || c->win_sub == 4)) ? 1 : 0;
    if (use_win && c->win_unroll == 8 && c->win_minb == 4 && (c->win_sub == 2 || c->win_sub == 4)) k_win = win_variant(c->win_sub, use_stab);
    if (use_win && c->win_stab && stab_ok && !use_stab) {           // other shapes of the shared-memory variant (A/B runs through the environment)
        switch (c->win_unroll * 100 + c->win_minb * 10 + c->win_sub) {
            case 848: k_win = k_full_contacts_win<8, 4, 8, LAW7_M>; use_stab = 1; break;
            case 834: k_win = k_full_contacts_win<8, 3, 4, LAW7_M>; use_stab = 1; break;
            case 854: k_win = k_full_contacts_win<8, 5, 4, LAW7_M>; use_stab = 1; break;
            case 864: k_win = k_full_contacts_win<8, 6, 4, LAW7_M>; use_stab = 1; break;
            case 444: k_win = k_full_contacts_win<4, 4, 4, LAW7_M>; use_stab = 1; break;
            case 464: k_win = k_full_contacts_win<4, 6, 4, LAW7_M>; use_stab = 1; break;
            default: break;
        }
        if (use_stab) {
            int fc_blocks = 0;
            CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fc_blocks, k_win, 256, 0));
            g1 = (int)std::min<long long>(std::min(ps, c->n_sm * std::max(1, fc_blocks)), ((long long)c->n_items + 7) / 8);
        }
    }
    const bool cached = !p_override && c->band_slot == slot && c->band_age < GRAAL_BAND_RESYNC;
    auto enqueue = [&]() -> int {
        // d_out = -(lf_total + G0) [+ log g * sum(ob)]   then accumulate the device sums
        const double init = -(c->lf_total + g0) + (uniform ? lg_uniform * c->ob_total : 0.0);
        k_set_double<<<1, 1, 0, st>>>(d_out, init); CHECK_LAUNCH(c);
        if (!(use_win && g1 > 0) && !cached) {      // contig offsets of the position order (the windowed pass left them behind)
            CUDA_OK(cudaMemsetAsync(c->cont_len, 0, (size_t)c->cap * sizeof(int), st));
            k_contig_lengths<<<nblk(n, 256), 256, 0, st>>>(s, ld, n, c->cap, c->cont_len); CHECK_LAUNCH(c);
            size_t tb = c->cub_tmp_bytes;
            CUDA_OK(cub::DeviceScan::ExclusiveSum(c->cub_tmp, tb, c->cont_len, c->cont_off, c->cap, st)); c->launches += 2;
        }
        if (use_win && g1 > 0) {
            int* bad = c->d_ints + 4;
            const FastLaw fl = win_law(use_stab);
            c->prof.begin(GRAAL_K_FULL_CONTACTS, st);
            k_win<<<g1, 256, 0, st>>>(c->items, c->item_hdr, c->n_items, c->contacts, c->cm_base, bad, c->W, c->geo_base, fl, p, lg_uniform, c->partials);
            CHECK_LAUNCH(c);
            c->prof.end(GRAAL_K_FULL_CONTACTS, st);
            k_reduce_partials<<<1, 256, 0, st>>>(c->partials, g1, 0, 1.0, d_out, 1); CHECK_LAUNCH(c);
        } else if (g1 > 0) {
            c->prof.begin(GRAAL_K_FULL_CONTACTS, st);
            if (use_smem)
                k_full_contacts_uniform<true><<<g1, FCS_THREADS, fcs_smem, st>>>(c->rowptr, c->contacts, c->E, c->W, c->group_row, c->n_groups, c->geo_base,
                                                                                c->cid16_base, c->mid32_base, p, lg_uniform, c->partials);
            else if (uniform && p.mode == 2)
                k_full_contacts_direct<<<g1, 256, 0, st>>>(c->rowptr, c->contacts, c->E, c->group_row, c->n_groups, c->geo_base,
                                                          c->cm_base, p, lg_uniform, c->partials);
            else if (uniform)
                k_full_contacts_uniform<false><<<g1, 256, 0, st>>>(c->rowptr, c->contacts, c->E, c->W, c->group_row, c->n_groups, c->geo_base,
                                                                  c->cid16_base, c->mid32_base, p, lg_uniform, c->partials);
            else
                k_full_contacts<<<g1, 256, 0, st>>>(c->rowptr, c->contacts, c->E, c->W, c->geo_base, p, c->partials);
            CHECK_LAUNCH(c);
            c->prof.end(GRAAL_K_FULL_CONTACTS, st);
            k_reduce_partials<<<1, 256, 0, st>>>(c->partials, g1, 0, 1.0, d_out, 1); CHECK_LAUNCH(c);
        }
        const int g2 = std::min(ps, nblk(n, 8));
        c->prof.begin(GRAAL_K_FULL_BAND, st);
        if (cached) {
            const int g3 = std::min(ps, nblk(n, 256));
            k_band_diag<<<g3, 256, 0, st>>>(s, ld, n, c->lv, c->geo_base, p, c->partials + (size_t)1 * ps); CHECK_LAUNCH(c);
            k_reduce_partials<<<1, 256, 0, st>>>(c->partials + (size_t)1 * ps, g3, 0, -1.0, d_out, 1); CHECK_LAUNCH(c);
            k_reduce_partials<<<1, 32, 0, st>>>(c->d_scalars + 40, 1, 0, -1.0, d_out, 1); CHECK_LAUNCH(c);
        } else {
            // position order of the bins, contig by contig (already filled by the windowed pass)
            if (!(use_win && g1 > 0)) { k_order_fill<<<nblk(n, 256), 256, 0, st>>>(s, ld, n, c->cap, c->cont_off, c->order); CHECK_LAUNCH(c); }
            k_band<<<g2, 256, 0, st>>>(c->order, n, s, ld, c->lv, c->geo_base, p, c->partials, ps); CHECK_LAUNCH(c);
            double* cross = p_override ? c->d_scalars + 41 : c->d_scalars + 40;
            k_reduce_partials<<<1, 256, 0, st>>>(c->partials, g2, 0, 1.0, cross, 0); CHECK_LAUNCH(c);
            k_reduce_partials<<<1, 32, 0, st>>>(cross, 1, 0, -1.0, d_out, 1); CHECK_LAUNCH(c);
            k_reduce_partials<<<1, 256, 0, st>>>(c->partials + (size_t)1 * ps, g2, 0, -1.0, d_out, 1); CHECK_LAUNCH(c);
        }
        c->prof.end(GRAAL_K_FULL_BAND, st);
        if (c->n_rep > 0) {     // every pixel touching a duplicated data bin
            const int gr = (int)std::min<long long>(ps, ((long long)c->n_rep * c->N + 127) / 128);
            k_repeat_pixels<false><<<dim3(gr, 1), 128, 0, st>>>(s, 0, ld, c->lv, c->collector, reinterpret_cast<const int2*>(c->dispatcher), c->rowptr, c->contacts,
                                                               c->d_rep_bins, c->n_rep, nullptr, p, c->partials + (size_t)3 * ps, ps); CHECK_LAUNCH(c);
            k_reduce_partials<<<1, 256, 0, st>>>(c->partials + (size_t)3 * ps, gr, 0, 1.0, d_out, 1); CHECK_LAUNCH(c);
        }
        if (c->n_quirky > 0) {
            k_quirk<<<dim3(c->n_quirky, 1), 256, 0, st>>>(c->d_quirky, c->n_quirky, s, ld, n, 0, c->lv, nullptr, nullptr, p,
                                                          c->partials + (size_t)2 * ps, ps); CHECK_LAUNCH(c);
            k_reduce_partials<<<1, 256, 0, st>>>(c->partials + (size_t)2 * ps, c->n_quirky, 0, -1.0, d_out, 1); CHECK_LAUNCH(c);
        }
        return 0;
    };
    // current parameters: the sequence only depends on (slot, output, cached band total) -> one graph launch
    if (p_override) rc = enqueue();
    else rc = run_graphed(c, cached ? c->g_full_cached : c->g_full, slot, (long long)(uintptr_t)d_out, 0, enqueue);
    if (rc) return rc;
    if (cached) c->band_age++;
    else if (!p_override) { c->band_slot = slot; c->band_age = 0; }
    return 0;
}

// union windows in the delta contact pass: uniform-accu level (every far pair has the same value) without repeats
static bool delta_uni_ok(const graal_ctx* c) {
    return c->delta_uni && c->p.nd == 1 && c->n_rep == 0 && c->n_quirky == 0 && c->E > 0 && c->p.d_max > 0.0f;
}
// the base slot's geometry (and, for the windowed delta pass, its row windows) must be current before this runs on `st`
// genome distance of the 13 candidates of a proposal (graal_dist_candidates), queued on a side stream of the lane beside the
// delta passes (it only needs the candidate slots) instead of behind them
struct DistArgs { const int32_t* init_prev = nullptr; const int32_t* init_next = nullptr; const int32_t* init_orientable = nullptr;
                  const uint8_t* skip = nullptr; double* d_out = nullptr; };
static int delta_loglik_impl(graal_ctx* c, Lane& L, cudaStream_t st, int base_slot, int first_cand_slot, int n_cand, int id_fA, int id_fB, int max_id,
                            unsigned skip, double* d_out, double* d_band, int copy_to = 0, unsigned pair_mask = 0u, const DistArgs* dist = nullptr) {
    const int n = c->n_new, ld = c->ld;
    const Params p = c->p;
    const bool uni = delta_uni_ok(c) && windows_current(c, base_slot, p.d_max);
    int* base = slot_ptr(c, base_slot);
    int* cand0 = slot_ptr(c, first_cand_slot);
    int* meta = L.ints + 8;
    int* piece_len = L.ints + 16;
    const int ps = c->partial_stride;
    int* rng = L.ints + 160;                       // [0,1] base range, [2 + 2k, 3 + 2k] candidate k
    k_delta_setup<<<nblk(n, 256), 256, 0, st>>>(base, ld, id_fA, id_fB, c->d_ints + 0, max_id, meta, n, c->W, L.sub_index, L.chmask, piece_len, rng, L.chmask2); CHECK_LAUNCH(c);
    const int gu = std::min(c->n_sm * 2, nblk(n, 256));
    if (c->prof.on) { k_u_stats<<<gu, 256, 0, st>>>(L.sub_index, meta, c->lv, base, ld, c->rowptr, c->d_counters); CHECK_LAUNCH(c); }
    k_cand_geometry<<<dim3(gu, n_cand), 256, 0, st>>>(cand0, slot_stride(c), ld, c->lv, L.sub_index, meta, L.geo_cand, (size_t)c->W, piece_len,
                                                     c->geo_base, L.chmask, skip, meta + 6, L.cand_ordrec, n); CHECK_LAUNCH(c);
    const int n_rows = pair_mask ? N_BAND_ROWS : n_cand;          // band rows: candidates (+ the partner rows of the paired ones)
    k_cand_order<<<dim3(gu, n_cand + 1 + (pair_mask ? 3 : 0)), 256, 0, st>>>(cand0, slot_stride(c), ld, L.sub_index, meta, piece_len, n, skip,
                                                      c->lv, L.chmask, L.geo_cand, (size_t)c->W, L.cand_ordrec, L.cand_ordb, rng + 2,
                                                      n_cand, base, c->geo_base, L.base_ordrec, L.base_ordb, L.base_ordc, rng,
                                                      pair_mask, L.chmask2, meta + 6, c->row_hdr, uni ? L.uwin : nullptr); CHECK_LAUNCH(c);
    const int nblk32 = n / 32 + 1;
    if (uni) {      // union over the candidates of the window of every bin of U
        k_cand_hull_blocks<<<dim3(gu, n_cand), 256, 0, st>>>(L.cand_ordrec, n, meta, L.cand_blk, nblk32, skip); CHECK_LAUNCH(c);
        const float reach = p.d_max * 1.0001f + c->max_bin_kb + 0.1f;
        k_cand_windows<<<dim3(gu, n_cand), 256, 0, st>>>(L.cand_ordrec, n, meta, piece_len, L.cand_blk, nblk32, skip, reach, L.uwin); CHECK_LAUNCH(c);
    }
    // grid-stride kernels over the (device-side) size of U: a few CTAs per SM, not one warp per bin of the level
    const int gw = std::min(ps, std::max(1, std::min(nblk(n, 1), c->n_sm * c->delta_minb)));      // one resident wave (128 registers: 2 CTAs per SM)
    // (band: 2 CTAs per SM and one warp per x measured best with three proposals in flight: small grids share the SMs)
    const int gb = std::min(ps, std::max(1, std::min(nblk(n, 4), c->n_sm * 2)));
    double* p_contacts = L.partials, *p_cand = L.partials + (size_t)16 * ps, *p_base = L.partials + (size_t)32 * ps;
    // the three passes are independent: contacts stay on `st`, the two band passes go to the lane's side streams
    // (one stream when the per-kernel profiler is on, or without side streams)
    const bool fork = c->fork_passes && !c->prof.on && L.side[0] && L.side[1];
    cudaStream_t s_cand = fork ? L.side[0] : st, s_base = fork ? L.side[1] : st;
    if (fork) {
        CUDA_OK(cudaEventRecord(L.ev_ready, st));
        CUDA_OK(cudaStreamWaitEvent(s_cand, L.ev_ready, 0));
        CUDA_OK(cudaStreamWaitEvent(s_base, L.ev_ready, 0));
    }
    if (dist && dist->d_out) {      // on the stream of the (short) base band pass
        const int gd = std::min(ps, std::max(1, nblk(n, 256)));
        k_dist_genome<<<dim3(gd, n_cand), 256, 0, s_base>>>(cand0, slot_stride(c), ld, n, dist->init_prev, dist->init_next, dist->init_orientable, dist->skip,
                                                           L.dist_partials, ps); CHECK_LAUNCH(c);
        k_reduce_partials<<<n_cand, 256, 0, s_base>>>(L.dist_partials, gd, ps, 1.0, dist->d_out, 0); CHECK_LAUNCH(c);
    }
    // contacts: sum over changed contacts of ob * (ln ex_k - ln ex_0): new terms per candidate, old terms once
    c->prof.begin(GRAAL_K_DELTA_CONTACTS, st);
    FastLaw flc; flc.tab = p.t_lnfu; flc.smin_bits = LAW_SMIN_BITS; flc.span = p.fu_span; flc.zlo = p.fu_zlo; flc.zspan = p.fu_zspan;
    const bool rel = c->delta_rel && p.nd == 1 && p.mode == 2 && p.fu_ok && c->n_rep == 0;       // relative contact terms (uniform level, tabulated monotone law)
    const double lg = rel ? log((double)g_clamp(c->accu_hist[0].first * c->accu_hist[0].first, p.v_inter, p.nfpb)) : 0.0;
#define DC_LAUNCH(U_, B_) k_delta_contacts_rows<U_, B_><<<dim3(gw, 1), 256, 0, st>>>(c->rowptr, c->contacts, c->lv, L.sub_index, meta, c->geo_base, L.geo_cand, (size_t)c->W, \
                                                                L.chmask, L.chmask2, pair_mask, p, p_contacts, ps, uni ? L.uwin : nullptr, flc, lg, c->delta_split)
    if (rel) { if (c->delta_minb == 4) DC_LAUNCH(true, 4); else if (c->delta_minb == 3) DC_LAUNCH(true, 3); else DC_LAUNCH(true, 2); }
    else     { if (c->delta_minb == 4) DC_LAUNCH(false, 4); else if (c->delta_minb == 3) DC_LAUNCH(false, 3); else DC_LAUNCH(false, 2); }
#undef DC_LAUNCH
    CHECK_LAUNCH(c);
    c->prof.end(GRAAL_K_DELTA_CONTACTS, st);
    // band mass: d_band[k] = B_U(S_k) - B_U(S_0) over changed pairs; enters the delta with a minus sign
    c->prof.begin(GRAAL_K_DELTA_BAND, st);
    const bool fast_band = c->band_fast && p.nd == 1 && p.mode == 2 && p.fu_ok;
    FastBand fbd; fbd.tab = p.t_fu; fbd.smin_bits = LAW_SMIN_BITS; fbd.span = p.fu_span; fbd.zlo = p.fu_zlo; fbd.zspan = p.fu_zspan;
    { union { float f; unsigned u; } dm; dm.f = p.d_max; fbd.dmax_bits = dm.u; }
    if (fast_band) k_band_delta_fast<false, 1><<<dim3(gb, n_rows), 256, 0, s_cand>>>(L.cand_ordrec, L.cand_ordb, nullptr, n, meta + 4, rng + 2, L.geo_cand, (size_t)c->W, skip, fbd, p,
                                                                              p_cand, ps, c->band_split);
    else if (p.nd == 1) k_band_delta<false, 1, true><<<dim3(gb, n_rows), 256, 0, s_cand>>>(L.cand_ordrec, L.cand_ordb, nullptr, n, meta + 4, rng + 2, L.geo_cand, (size_t)c->W, skip, p,
                                                                              p_cand, ps);
    else k_band_delta<false, 1, false><<<dim3(gb, n_rows), 256, 0, s_cand>>>(L.cand_ordrec, L.cand_ordb, nullptr, n, meta + 4, rng + 2, L.geo_cand, (size_t)c->W, skip, p,
                                                                     p_cand, ps);
    CHECK_LAUNCH(c);
    if (fast_band) k_band_delta_fast<true, 4><<<dim3(gb, 1), 256, 0, s_base>>>(L.base_ordrec, L.base_ordb, L.base_ordc, n, meta + 4, rng, c->geo_base, 0, skip, fbd, p,
                                                                        p_base, ps, c->band_split);
    else if (p.nd == 1) k_band_delta<true, 4, true><<<dim3(gb, 1), 256, 0, s_base>>>(L.base_ordrec, L.base_ordb, L.base_ordc, n, meta + 4, rng, c->geo_base, 0, skip, p,
                                                                        p_base, ps);
    else k_band_delta<true, 4, false><<<dim3(gb, 1), 256, 0, s_base>>>(L.base_ordrec, L.base_ordb, L.base_ordc, n, meta + 4, rng, c->geo_base, 0, skip, p,
                                                               p_base, ps);
    CHECK_LAUNCH(c);
    c->prof.end(GRAAL_K_DELTA_BAND, st);
    if (fork) {
        CUDA_OK(cudaEventRecord(L.ev_side[0], s_cand)); CUDA_OK(cudaEventRecord(L.ev_side[1], s_base));
        CUDA_OK(cudaStreamWaitEvent(st, L.ev_side[0], 0)); CUDA_OK(cudaStreamWaitEvent(st, L.ev_side[1], 0));
    }
    // second stage of the three reductions, out = contacts - (cand band - base band); candidate 8 (skipped) copies candidate 0
    k_finish_delta<<<n_cand, 256, 0, st>>>(p_contacts, gw, p_cand, p_base, gb, ps, d_out, d_band, copy_to, pair_mask); CHECK_LAUNCH(c);
    if (c->n_rep > 0) {     // ranges 2-4: pixels of the duplicated bins that have a copy in U, new minus old
        CUDA_OK(cudaMemsetAsync(L.rep_in_u, 0, (size_t)c->N, st));
        k_mark_rep_in_u<<<gu, 256, 0, st>>>(base, ld, L.sub_index, meta, c->lv, L.rep_in_u); CHECK_LAUNCH(c);
        const int gr = (int)std::min<long long>(ps, ((long long)c->n_rep * c->N + 127) / 128);
        k_repeat_pixels<true><<<dim3(gr, n_cand), 128, 0, st>>>(cand0, slot_stride(c), ld, c->lv, c->collector, reinterpret_cast<const int2*>(c->dispatcher),
                                                               c->rowptr, c->contacts, c->d_rep_bins, c->n_rep, L.rep_in_u, p, L.partials, ps); CHECK_LAUNCH(c);
        k_reduce_partials<<<n_cand, 256, 0, st>>>(L.partials, gr, ps, 1.0, d_out, 1); CHECK_LAUNCH(c);
        k_repeat_pixels<true><<<dim3(gr, 1), 128, 0, st>>>(base, 0, ld, c->lv, c->collector, reinterpret_cast<const int2*>(c->dispatcher),
                                                          c->rowptr, c->contacts, c->d_rep_bins, c->n_rep, L.rep_in_u, p, L.partials + (size_t)13 * ps, ps); CHECK_LAUNCH(c);
        for (int k = 0; k < n_cand; k++) {
            k_reduce_partials<<<1, 256, 0, st>>>(L.partials + (size_t)13 * ps, gr, 0, -1.0, d_out + k, 1); CHECK_LAUNCH(c);
        }
    }
    if (c->n_quirky > 0) {
        // quirk mass: - [Q_U(S_k) - Q_U(S_0)]
        k_quirk<<<dim3(c->n_quirky, n_cand), 256, 0, st>>>(c->d_quirky, c->n_quirky, cand0, ld, n, slot_stride(c), c->lv, L.sub_index, meta + 4, p,
                                                          L.partials, ps); CHECK_LAUNCH(c);
        k_reduce_partials<<<n_cand, 256, 0, st>>>(L.partials, c->n_quirky, ps, -1.0, d_out, 1); CHECK_LAUNCH(c);
        // the base slot's term is the same for every candidate: computed once, added to all
        k_quirk<<<dim3(c->n_quirky, 1), 256, 0, st>>>(c->d_quirky, c->n_quirky, base, ld, n, 0, c->lv, L.sub_index, meta + 4, p,
                                                     L.partials + (size_t)13 * ps, ps); CHECK_LAUNCH(c);
        for (int k = 0; k < n_cand; k++) {
            k_reduce_partials<<<1, 256, 0, st>>>(L.partials + (size_t)13 * ps, c->n_quirky, 0, 1.0, d_out + k, 1); CHECK_LAUNCH(c);
        }
    }
    return 0;
}

int graal_delta_loglik(graal_ctx* c, int base_slot, int first_cand_slot, int n_cand, int id_fA, int id_fB, int max_id, double* d_out) {
    NEED_STATE(c); NEED_SLOT(c, base_slot); NEED_SLOT(c, first_cand_slot);
    if (n_cand < 1 || n_cand > GRAAL_N_CANDIDATES) return set_err(-1, "n_cand must be 1..13");
    NEED_SLOT(c, first_cand_slot + n_cand - 1);
    if (!d_out) return set_err(-1, "null output");
    if (!c->have_params) return set_err(-1, "parameters not set");
    if (id_fA < 0 || id_fA >= c->n_new || id_fB < 0 || id_fB >= c->n_new) return set_err(-1, "bin id out of range");
    int rc = ensure_base_geometry(c, base_slot); if (rc) return rc;
    if (delta_uni_ok(c)) { rc = ensure_base_windows(c, base_slot, c->p.d_max); if (rc) return rc; }
    return delta_loglik_impl(c, c->lanes[0], c->stream, base_slot, first_cand_slot, n_cand, id_fA, id_fB, max_id, 0u, d_out, c->d_scalars + 16);
}

static int score_proposal_impl(graal_ctx* c, int base_slot, int first_cand_slot, int id_fA, int id_fB, int max_id, int proposal_index, double* d_out, const DistArgs* dist) {
    if (!c || !c->slots) return set_err(-1, "state not bound");
    CUDA_OK(cudaSetDevice(c->device));
    if (proposal_index < 0 || proposal_index >= 16) return set_err(-1, "proposal index must be 0..15");
    if (!d_out) return set_err(-1, "null output");
    if (!c->have_params) return set_err(-1, "parameters not set");
    NEED_SLOT(c, base_slot); NEED_SLOT(c, first_cand_slot); NEED_SLOT(c, first_cand_slot + GRAAL_N_CANDIDATES - 1);
    const int n = c->n_new;
    if (id_fA < 0 || id_fA >= n || id_fB < 0 || id_fB >= n) return set_err(-1, "bin id out of range");
    if (base_slot >= first_cand_slot && base_slot < first_cand_slot + GRAAL_N_CANDIDATES) return set_err(-1, "source slot inside the destination range");
    // Lane of this proposal.  It runs concurrently with the proposals pending on the other lanes unless its
    // candidate slots overlap theirs (a caller with a single set of 13 candidate slots scores serially) or the
    // per-kernel profiler is on (its brackets assume one stream).
    Lane& L = c->lanes[proposal_index % c->n_lanes];
    bool serial = c->prof.on || c->n_lanes == 1;
    for (int l = 0; l < c->n_lanes && !serial; l++) {
        const Lane& o = c->lanes[l];
        if (&o != &L && o.pending && first_cand_slot < o.cand_first + GRAAL_N_CANDIDATES && o.cand_first < first_cand_slot + GRAAL_N_CANDIDATES) serial = true;
    }
    int rc;
    if (serial) { rc = join_lanes(c); if (rc) return rc; }
    rc = ensure_base_geometry(c, base_slot); if (rc) return rc;
    if (delta_uni_ok(c)) { rc = ensure_base_windows(c, base_slot, c->p.d_max); if (rc) return rc; }
    cudaStream_t st = c->stream;
    if (!serial) {
        CUDA_OK(cudaEventRecord(c->ev_fork, c->stream));
        CUDA_OK(cudaStreamWaitEvent(L.st, c->ev_fork, 0));
        st = L.st;
    }
    for (int k = 0; k < GRAAL_N_CANDIDATES; k++) {
        if (c->geo_base_slot == first_cand_slot + k) c->geo_base_slot = -1;
        if (c->first_idx_slot == first_cand_slot + k) c->first_idx_slot = -1;
        if (c->band_slot == first_cand_slot + k) c->band_slot = -1;
    }
    // unique bins: swap_activity is the identity on the popped-out structure, candidate 8 == candidate 0 (Q7)
    const unsigned skip = (id_fA < c->N) ? (1u << 8) : 0u;      // a repeat copy (frag >= N) really toggles its activity
    double* d_band = c->band_hist + (size_t)proposal_index * GRAAL_N_CANDIDATES;
    // the launch sequence of one proposal (candidates, deltas, copy of candidate 8)
    auto enqueue = [&]() -> int {
        c->prof.begin(GRAAL_K_BUILD, st);
        k_build_candidates<<<nblk(n, 128), 128, 0, st>>>(slot_ptr(c, base_slot), slot_ptr(c, first_cand_slot), slot_stride(c), c->ld, n,
                                                        id_fA, id_fB, c->d_ints + 0, max_id, 0x1FFFu);
        CHECK_LAUNCH(c);
        c->prof.end(GRAAL_K_BUILD, st);
        return delta_loglik_impl(c, L, st, base_slot, first_cand_slot, GRAAL_N_CANDIDATES, id_fA, id_fB, max_id, skip, d_out, d_band, skip ? 8 : 0,
                                 c->pairing ? PAIR_MASK : 0u, dist);
    };
    ProposalGraph& G = c->graphs[proposal_index];
    if (!c->use_graphs || c->prof.on) { rc = enqueue(); if (rc) return rc; }
    else {
        const bool hit = G.exec && G.version == c->version && G.base_slot == base_slot && G.first_cand == first_cand_slot &&
                         G.d_out == d_out && G.skip == skip && G.st == st && G.d_dist == (dist ? dist->d_out : nullptr);
        if (!hit) {                                             // capture the sequence for this proposal index
            G.reset();
            const int64_t before = c->launches;
            CUDA_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            rc = enqueue();
            cudaGraph_t graph = nullptr;
            const cudaError_t e_end = cudaStreamEndCapture(st, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e_end != cudaSuccess) return set_err(-2, "stream capture of a proposal failed: %s", cudaGetErrorString(e_end));
            G.graph = graph; G.n_launches = (int)(c->launches - before); c->launches = before;
            size_t nn = 0;
            CUDA_OK(cudaGraphGetNodes(graph, nullptr, &nn));
            std::vector<cudaGraphNode_t> nodes(nn);
            CUDA_OK(cudaGraphGetNodes(graph, nodes.data(), &nn));
            for (size_t i = 0; i < nn; i++) {
                cudaGraphNodeType t;
                CUDA_OK(cudaGraphNodeGetType(nodes[i], &t));
                if (t != cudaGraphNodeTypeKernel) continue;
                cudaKernelNodeParams kp;
                CUDA_OK(cudaGraphKernelNodeGetParams(nodes[i], &kp));
                if (kp.func == (void*)k_build_candidates) { G.n_build = nodes[i]; G.kp_build = kp; }
                else if (kp.func == (void*)k_delta_setup) { G.n_setup = nodes[i]; G.kp_setup = kp; }
            }
            if (!G.n_build || !G.n_setup) { G.reset(); return set_err(-2, "proposal graph: parameter nodes not found"); }
            CUDA_OK(cudaGraphInstantiate(&G.exec, graph, 0));
            G.version = c->version; G.base_slot = base_slot; G.first_cand = first_cand_slot; G.d_out = d_out; G.skip = skip; G.st = st;
            G.d_dist = dist ? dist->d_out : nullptr;
        } else {                                                // same sequence, new (id_fA, id_fB, max_id)
            const int* a_src = slot_ptr(c, base_slot); int* a_dst = slot_ptr(c, first_cand_slot); size_t a_stride = slot_stride(c);
            int a_ld = c->ld, a_n = n, a_fA = id_fA, a_fB = id_fB, a_max = max_id; const int* a_dmax = c->d_ints + 0; unsigned a_mask = 0x1FFFu;
            void* args_build[] = {&a_src, &a_dst, &a_stride, &a_ld, &a_n, &a_fA, &a_fB, &a_dmax, &a_max, &a_mask};
            cudaKernelNodeParams kp = G.kp_build; kp.kernelParams = args_build; kp.extra = nullptr;
            CUDA_OK(cudaGraphExecKernelNodeSetParams(G.exec, G.n_build, &kp));
            int* a_meta = L.ints + 8; int a_W = c->W; int* a_sub = L.sub_index; unsigned* a_chm = L.chmask; int* a_pl = L.ints + 16; int* a_rng = L.ints + 160; unsigned* a_chm2 = L.chmask2;
            void* args_setup[] = {&a_src, &a_ld, &a_fA, &a_fB, &a_dmax, &a_max, &a_meta, &a_n, &a_W, &a_sub, &a_chm, &a_pl, &a_rng, &a_chm2};
            kp = G.kp_setup; kp.kernelParams = args_setup; kp.extra = nullptr;
            CUDA_OK(cudaGraphExecKernelNodeSetParams(G.exec, G.n_setup, &kp));
        }
        CUDA_OK(cudaGraphLaunch(G.exec, st));
        c->launches += G.n_launches;
    }
    if (!serial) {
        CUDA_OK(cudaEventRecord(L.done, L.st));
        L.pending = true; L.cand_first = first_cand_slot;
    }
    return 0;
}

int graal_score_proposal(graal_ctx* c, int base_slot, int first_cand_slot, int id_fA, int id_fB, int max_id, int proposal_index, double* d_out) {
    return score_proposal_impl(c, base_slot, first_cand_slot, id_fA, id_fB, max_id, proposal_index, d_out, nullptr);
}

int graal_score_step(graal_ctx* c, int base_slot, int first_cand_slot, int id_fA, const int32_t* id_fB, int n_proposals, int max_id,
                     double* d_out, const int32_t* init_prev, const int32_t* init_next, const int32_t* init_orientable,
                     const uint8_t* skip, double* d_dist) {
    if (!c || !c->slots) return set_err(-1, "state not bound");
    if (!id_fB || n_proposals < 0 || n_proposals > 16) return set_err(-1, "need 0..16 proposals");
    // proposal x uses the candidate slots of lane x % n_lanes when the state block has them, else one shared range
    const bool per_lane = first_cand_slot + GRAAL_N_CANDIDATES * c->n_lanes <= c->n_slots;
    for (int x = 0; x < n_proposals; x++) {
        const int first = first_cand_slot + (per_lane ? GRAAL_N_CANDIDATES * (x % c->n_lanes) : 0);
        DistArgs da;
        if (d_dist) {
            if (!init_prev || !init_next || !init_orientable || !skip) return set_err(-1, "null argument");
            da.init_prev = init_prev; da.init_next = init_next; da.init_orientable = init_orientable; da.skip = skip;
            da.d_out = d_dist + (size_t)GRAAL_N_CANDIDATES * x;
        }
        // (the genome distance of the candidates rides on a side stream of the proposal's lane, inside its captured graph)
        int rc = score_proposal_impl(c, base_slot, first, id_fA, id_fB[x], max_id, x, d_out + (size_t)GRAAL_N_CANDIDATES * x, d_dist ? &da : nullptr); if (rc) return rc;
    }
    return 0;
}

int graal_fetch(graal_ctx* c, const void* d_src, void* h_dst, size_t bytes) {
    if (!c || !d_src || !h_dst) return set_err(-1, "null argument");
    CUDA_OK(cudaSetDevice(c->device));
    int rc = join_lanes(c); if (rc) return rc;
    CUDA_OK(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, c->stream));
    const bool pend = c->ncontigs_pending && c->h_ncontigs;
    if (pend) CUDA_OK(cudaMemcpyAsync(c->h_ncontigs, c->d_ints + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (pend) {     // contig ids of the relabelled slot are < n_contigs (+ 3 per candidate committed since): the fused prologue's bound
        c->contig_bound = *c->h_ncontigs + 3 * c->bound_commits; c->ncontigs_pending = false;
    }
    return 0;
}

int graal_commit_scored(graal_ctx* c, int base_slot, int first_cand_slot, int id_fA, int id_fB, int max_id, int mode, int proposal_index) {
    NEED_STATE(c);
    if (mode < 0 || mode >= GRAAL_N_CANDIDATES) return set_err(-1, "mode must be 0..12");
    if (proposal_index >= 16) return set_err(-1, "proposal index must be < 16");
    const unsigned mask = (mode < 9) ? (1u << mode) : 0x1E00u;            // test_copy_struct (cuda_lib_gl.py:1156-1183)
    const int keep = c->band_slot;
    int rc = graal_build_candidates(c, base_slot, first_cand_slot, id_fA, id_fB, max_id, mask); if (rc) return rc;
    rc = graal_commit(c, base_slot, first_cand_slot + mode); if (rc) return rc;
    if (proposal_index >= 0 && keep == base_slot) {       // keep the cached band total in step with the committed candidate
        k_add_selected<<<1, 1, 0, c->stream>>>(c->d_scalars + 40, c->band_hist + (size_t)proposal_index * GRAAL_N_CANDIDATES, mode); CHECK_LAUNCH(c);
        c->band_slot = base_slot;
    }
    return 0;
}

int graal_state_stats(graal_ctx* c, int slot, double* d_out) {
    NEED_STATE(c); NEED_SLOT(c, slot);
    if (!d_out) return set_err(-1, "null output");
    const int n = c->n_new, ld = c->ld, cap = c->cap;
    int* s = slot_ptr(c, slot);
    cudaStream_t st = c->stream;
    auto enqueue = [&]() -> int {
        k_init_stats<<<1, 1, 0, st>>>(c->d_stats); CHECK_LAUNCH(c);
        k_stats<<<std::min(nblk(n, 256), c->n_sm * 4), 256, 0, st>>>(s, ld, n, c->d_stats); CHECK_LAUNCH(c);
        k_fill_int<<<nblk(cap, 256), 256, 0, st>>>(c->first_idx, cap, INT_MAX); CHECK_LAUNCH(c);
        k_first_index<<<nblk(n, 256), 256, 0, st>>>(s + F_ID_C * ld, n, cap, c->first_idx, c->d_ints + 2); CHECK_LAUNCH(c);
        k_set_int<<<1, 1, 0, st>>>(c->d_ints + 3, 0); CHECK_LAUNCH(c);
        k_count_contigs<<<std::min(nblk(cap, 256), c->n_sm * 4), 256, 0, st>>>(c->first_idx, cap, c->d_ints + 3); CHECK_LAUNCH(c);
        k_stats_final<<<1, 1, 0, st>>>(c->d_stats, c->d_ints + 3, d_out); CHECK_LAUNCH(c);
        return 0;
    };
    { const int rc = run_graphed(c, c->g_stats, slot, (long long)(uintptr_t)d_out, 0, enqueue); if (rc) return rc; }
    c->first_idx_slot = slot; c->first_clean = false; c->stats_clean = false;
    return 0;
}

int graal_dist_genome(graal_ctx* c, int slot, const int32_t* init_prev, const int32_t* init_next, const int32_t* init_orientable,
                      const uint8_t* skip, double* d_out) {
    NEED_STATE(c); NEED_SLOT(c, slot);
    if (!init_prev || !init_next || !init_orientable || !skip || !d_out) return set_err(-1, "null argument");
    const int n = c->n_new;
    const int g = std::min(c->partial_stride, std::max(1, nblk(n, 256)));
    double* part = c->partials + (size_t)15 * c->partial_stride;
    k_dist_genome<<<g, 256, 0, c->stream>>>(slot_ptr(c, slot), 0, c->ld, n, init_prev, init_next, init_orientable, skip, part, 0); CHECK_LAUNCH(c);
    k_reduce_partials<<<1, 256, 0, c->stream>>>(part, g, 0, 1.0, d_out, 0); CHECK_LAUNCH(c);
    return 0;
}

int graal_dist_candidates(graal_ctx* c, int first_cand_slot, int n_cand, int proposal_index, const int32_t* init_prev, const int32_t* init_next,
                          const int32_t* init_orientable, const uint8_t* skip, double* d_out) {
    if (!c || !c->slots) return set_err(-1, "state not bound");
    CUDA_OK(cudaSetDevice(c->device));
    if (n_cand < 1 || n_cand > GRAAL_N_CANDIDATES) return set_err(-1, "n_cand must be 1..13");
    NEED_SLOT(c, first_cand_slot); NEED_SLOT(c, first_cand_slot + n_cand - 1);
    if (!init_prev || !init_next || !init_orientable || !skip || !d_out) return set_err(-1, "null argument");
    if (proposal_index >= 16) return set_err(-1, "proposal index must be < 16");
    // on the lane that is scoring these candidates, right behind its kernels; otherwise on the context stream
    Lane* L = (proposal_index >= 0) ? &c->lanes[proposal_index % c->n_lanes] : nullptr;
    cudaStream_t st = c->stream; double* part = c->partials; 
    if (L && L->pending && L->cand_first == first_cand_slot) { st = L->st; part = L->partials; }
    else { int rc = join_lanes(c); if (rc) return rc; }
    const int n = c->n_new, ps = c->partial_stride;
    const int g = std::min(ps, std::max(1, nblk(n, 256)));
    k_dist_genome<<<dim3(g, n_cand), 256, 0, st>>>(slot_ptr(c, first_cand_slot), slot_stride(c), c->ld, n, init_prev, init_next, init_orientable, skip, part, ps); CHECK_LAUNCH(c);
    k_reduce_partials<<<n_cand, 256, 0, st>>>(part, g, ps, 1.0, d_out, 0); CHECK_LAUNCH(c);
    if (st != c->stream) CUDA_OK(cudaEventRecord(L->done, L->st));      // the join must cover this work too
    return 0;
}

// ------------------------------------------------------------------------------------------------
// COO -> row-segmented contact lists on the device (sampler.__init__, cuda_lib_gl.py:153-172: csr + csr.T, diagonal zeroed)
// ------------------------------------------------------------------------------------------------
__global__ void k_coo_keys(const int* __restrict__ r, const int* __restrict__ c, const float* __restrict__ v, long long n, long long W,
                           unsigned long long* __restrict__ keys, float* __restrict__ vals) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long a = r[i], b = c[i];
        const bool ok = a != b && a >= 0 && b >= 0 && a < W && b < W;
        keys[i] = ok ? (unsigned long long)(min(a, b) * W + max(a, b)) : ~0ull;       // diagonal / invalid entries sort last and are dropped
        vals[i] = ok ? v[i] : 0.0f;
    }
}
__global__ void k_coo_rows(const unsigned long long* __restrict__ keys, const float* __restrict__ vals, const int* __restrict__ n_unique, long long W,
                           unsigned char* __restrict__ keep, int* __restrict__ row_count) {
    const long long m = *n_unique;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        const bool k = keys[i] != ~0ull && vals[i] != 0.0f;
        keep[i] = k ? 1 : 0;
        if (k) atomicAdd(&row_count[(int)(keys[i] / (unsigned long long)W)], 1);
    }
}
__global__ void k_coo_emit(const unsigned long long* __restrict__ keys, const float* __restrict__ vals, const unsigned char* __restrict__ keep,
                           const long long* __restrict__ pos, const int* __restrict__ n_unique, long long W, int2* __restrict__ out) {
    const long long m = *n_unique;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x)
        if (keep[i]) out[pos[i]] = make_int2((int)(keys[i] % (unsigned long long)W), __float_as_int(vals[i]));
}

int graal_coo_to_lists(graal_ctx* c, const int32_t* d_rows, const int32_t* d_cols, const float* d_vals, int64_t n, int n_sub_frags,
                       int64_t* d_rowptr, void* d_contacts, int64_t* n_contacts_out) {
    if (!c || !d_rows || !d_cols || !d_vals || !d_rowptr || !d_contacts || !n_contacts_out) return set_err(-1, "null argument");
    if (n <= 0 || n >= (1ll << 31) || n_sub_frags <= 0) return set_err(-1, "bad sizes");
    CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const long long W = n_sub_frags;
    unsigned long long *k0 = nullptr, *k1 = nullptr, *ku = nullptr; float *v0 = nullptr, *v1 = nullptr, *vu = nullptr;
    int* d_nu = nullptr; int* d_cnt = nullptr; unsigned char* keep = nullptr; long long* pos = nullptr; void* tmp = nullptr;
    auto cleanup = [&]() { cudaFree(k0); cudaFree(k1); cudaFree(ku); cudaFree(v0); cudaFree(v1); cudaFree(vu); cudaFree(d_nu); cudaFree(d_cnt); cudaFree(keep); cudaFree(pos); cudaFree(tmp); };
    #define COO_OK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { cleanup(); return set_err(-2, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); } } while (0)
    COO_OK(cudaMalloc(&k0, n * 8)); COO_OK(cudaMalloc(&k1, n * 8)); COO_OK(cudaMalloc(&ku, n * 8));
    COO_OK(cudaMalloc(&v0, n * 4)); COO_OK(cudaMalloc(&v1, n * 4)); COO_OK(cudaMalloc(&vu, n * 4));
    COO_OK(cudaMalloc(&d_nu, 4)); COO_OK(cudaMalloc(&d_cnt, (size_t)(W + 1) * 4)); COO_OK(cudaMalloc(&keep, n)); COO_OK(cudaMalloc(&pos, n * 8));
    const int g = std::min<long long>((n + 255) / 256, (long long)c->n_sm * 16);
    k_coo_keys<<<g, 256, 0, st>>>(d_rows, d_cols, d_vals, n, W, k0, v0); c->launches++;
    size_t b1 = 0, b2 = 0, b3 = 0, b4 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, k0, k1, v0, v1, (int)n, 0, 64, st);
    cub::DeviceReduce::ReduceByKey(nullptr, b2, k1, ku, v1, vu, d_nu, cuda::std::plus<>(), (int)n, st);
    cub::DeviceScan::ExclusiveSum(nullptr, b3, (const unsigned char*)nullptr, (long long*)nullptr, (int)n, st);
    cub::DeviceScan::ExclusiveSum(nullptr, b4, (const int*)nullptr, (long long*)nullptr, (int)(W + 1), st);
    const size_t tb = std::max(std::max(b1, b2), std::max(b3, b4));
    COO_OK(cudaMalloc(&tmp, tb));
    size_t t = tb;
    // (invalid keys are all ones: they stay last whatever the number of key bits sorted)
    COO_OK(cub::DeviceRadixSort::SortPairs(tmp, t, k0, k1, v0, v1, (int)n, 0, 64, st)); t = tb;
    COO_OK(cub::DeviceReduce::ReduceByKey(tmp, t, k1, ku, v1, vu, d_nu, cuda::std::plus<>(), (int)n, st)); t = tb;      // duplicates summed (float32, in sorted order)
    COO_OK(cudaMemsetAsync(d_cnt, 0, (size_t)(W + 1) * 4, st));
    COO_OK(cudaMemsetAsync(keep, 0, n, st));
    k_coo_rows<<<g, 256, 0, st>>>(ku, vu, d_nu, W, keep, d_cnt); c->launches++;
    COO_OK(cub::DeviceScan::ExclusiveSum(tmp, t, keep, pos, (int)n, st)); t = tb;
    COO_OK(cub::DeviceScan::ExclusiveSum(tmp, t, d_cnt, reinterpret_cast<long long*>(d_rowptr), (int)(W + 1), st));
    k_coo_emit<<<g, 256, 0, st>>>(ku, vu, keep, pos, d_nu, W, reinterpret_cast<int2*>(d_contacts)); c->launches++;
    COO_OK(cudaMemcpyAsync(n_contacts_out, d_rowptr + W, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    COO_OK(cudaStreamSynchronize(st));
    #undef COO_OK
    cleanup();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Host arithmetic of the candidate draw (cuda_lib_gl.py:1899-1934), float64, in NumPy's operation order -- sums are NumPy's
// pairwise sums (numpy/core/src/umath/loops_utils.h: plain loop below 8 terms, eight interleaved partial sums up to
// 128 terms, halves rounded to multiples of 8 above) -- so that the weights are bit-identical to the NumPy statements they
// replace (tests/test_host.py compares them).  No device work: plain C called between the fetch and the commit of a step.
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// The candidate draw of step_max_likelihood (cuda_lib_gl.py:1899-1947) on the device: the arithmetic of
// graal_candidate_weights below (NumPy's operation order, pairwise sums included) by one warp, then the search of the uniform
// `u` (drawn by the host from the reference's stream, consumed only if more than one candidate is left) in the cumulative
// weights.  score = deltas + likelihood_t.  sel = {sample_out, n_ok, status (0 ok, 1 NaN / non-finite weights: the host path
// decides), id_f_sampled, op}; out_sub = the n_ok normalised weights (sub_score), out_score = {score of the drawn candidate,
// sample_out, n_ok, status} as doubles (one block for the host to fetch).
// ------------------------------------------------------------------------------------------------
// The step's output block straight into the caller's pinned host buffer (mapped: the device sees the same address), the contig
// count beside it, then -- behind a system-wide fence -- the sequence number graal_fetch_wait polls: no copy-engine operation
// between the draw and the commit kernels, no event.
__global__ void k_publish(const double* __restrict__ d_src, double* __restrict__ h_dst, int n_doubles, const int* __restrict__ d_ncontigs,
                          volatile int* __restrict__ h_ints, int with_ncontigs, int seq) {
    for (int i = threadIdx.x; i < n_doubles; i += blockDim.x) h_dst[i] = d_src[i];
    if (threadIdx.x == 0 && with_ncontigs) h_ints[0] = *d_ncontigs;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) { h_ints[8] = seq; __threadfence_system(); }
}
#define DRAW_MAX (16 * GRAAL_N_CANDIDATES)
struct NbList { int id[16]; };
__device__ double dev_np_sum128(const double* a, int n, int lane) {       // n <= 128, all lanes call, result in every lane
    double res;
    if (n < 8) { res = 0.0; for (int i = 0; i < n; i++) res += a[i]; return res; }
    double r = 0.0;
    const int n8 = n - (n % 8);
    if (lane < 8) { r = a[lane]; for (int i = 8; i < n8; i += 8) r += a[i + lane]; }
    const double r0 = __shfl_sync(0xffffffffu, r, 0), r1 = __shfl_sync(0xffffffffu, r, 1), r2 = __shfl_sync(0xffffffffu, r, 2), r3 = __shfl_sync(0xffffffffu, r, 3);
    const double r4 = __shfl_sync(0xffffffffu, r, 4), r5 = __shfl_sync(0xffffffffu, r, 5), r6 = __shfl_sync(0xffffffffu, r, 6), r7 = __shfl_sync(0xffffffffu, r, 7);
    res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (int i = n8; i < n; i++) res += a[i];
    return res;
}
__device__ double dev_np_pairwise_sum(const double* a, int n, int lane) {   // n <= 256
    if (n <= 128) return dev_np_sum128(a, n, lane);
    int n2 = n / 2; n2 -= n2 % 8;
    const double x = dev_np_sum128(a, n2, lane), y = dev_np_sum128(a + n2, n - n2, lane);
    return x + y;
}
__global__ void __launch_bounds__(32)
k_draw_candidates(const double* __restrict__ d_delta, const double* __restrict__ d_full, double likelihood_host, int use_host_likelihood,
                  int n, int n_tmp, double u, NbList nb, int* __restrict__ sel, double* __restrict__ out_sub, double* __restrict__ out_score) {
    __shared__ double score[DRAW_MAX], work[DRAW_MAX], cdf[DRAW_MAX];
    __shared__ int ids[DRAW_MAX];
    const int lane = threadIdx.x;
    const double lt = use_host_likelihood ? likelihood_host : *d_full;
    bool nan = false;
    for (int i = lane; i < n; i += 32) { const double v = d_delta[i] + lt; score[i] = v; if (v != v) nan = true; }
    __syncwarp();
    nan = __any_sync(0xffffffffu, nan);
    // argmax (first occurrence) and min
    int im = 0; double mn = score[0];
    if (lane == 0) for (int i = 1; i < n; i++) { if (score[i] > score[im]) im = i; if (score[i] < mn) mn = score[i]; }
    im = __shfl_sync(0xffffffffu, im, 0); mn = __shfl_sync(0xffffffffu, mn, 0);
    for (int i = lane; i < n; i += 32) {
        double w = score[i] - mn;
        if (i >= n_tmp && (i % n_tmp == 0 || i % n_tmp == 1)) w = 0.0;      // eject / flip of the later neighbours: duplicates
        work[i] = w;
    }
    __syncwarp();
    double mx = work[0];
    if (lane == 0) for (int i = 1; i < n; i++) if (work[i] > mx) mx = work[i];
    mx = __shfl_sync(0xffffffffu, mx, 0);
    const double shift = mx - 30.0;
    int n_ok = 0;
    for (int base = 0; base < n; base += 32) {                               // ordered compaction of the positive weights
        const int i = base + lane;
        double f = 0.0;
        if (i < n) { f = work[i] - shift; if (f < 0.0) f = 0.0; }
        const unsigned m = __ballot_sync(0xffffffffu, f > 0.0);
        if (f > 0.0) { const int k = n_ok + __popc(m & ((1u << lane) - 1u)); ids[k] = i; cdf[k] = f; }
        n_ok += __popc(m);
    }
    __syncwarp();
    int status = nan ? 1 : 0, sample = im;
    if (!nan && n_ok > 0) {
        double sum = dev_np_pairwise_sum(cdf, n_ok, lane);
        __syncwarp();
        for (int i = lane; i < n_ok; i += 32) cdf[i] = cdf[i] / sum;
        __syncwarp();
        sum = dev_np_pairwise_sum(cdf, n_ok, lane);
        __syncwarp();
        for (int i = lane; i < n_ok; i += 32) { cdf[i] = cdf[i] / sum; out_sub[i] = cdf[i]; }
        __syncwarp();
        if (lane == 0) { double run = 0.0; for (int i = 0; i < n_ok; i++) { run += cdf[i]; cdf[i] = run; } }
        __syncwarp();
        const double last = cdf[n_ok - 1];
        if (!(last - last == 0.0)) status = 1;
        else if (n_ok > 1) {
            int cnt = 0;
            for (int i = lane; i < n_ok; i += 32) if (cdf[i] / last <= u) cnt++;          // searchsorted(cdf / last, u, side='right')
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            sample = ids[min(cnt, n_ok - 1)];
        }
    }
    if (lane == 0) {
        sel[0] = sample; sel[1] = n_ok; sel[2] = status;
        sel[3] = nb.id[sample / n_tmp]; sel[4] = sample % n_tmp;
        out_score[0] = score[sample]; out_score[1] = (double)sample; out_score[2] = (double)n_ok; out_score[3] = (double)status;
    }
}

int graal_draw_candidates(graal_ctx* c, const double* d_delta, const double* d_full, double likelihood_host, int use_host_likelihood,
                          int n_proposals, const int32_t* id_fB, double u, int32_t* d_sel, double* d_sub_score, double* d_score) {
    NEED_STATE(c);
    if (!d_delta || !id_fB || !d_sel || !d_sub_score || !d_score || (!use_host_likelihood && !d_full)) return set_err(-1, "null argument");
    if (n_proposals < 1 || n_proposals > 16) return set_err(-1, "need 1..16 proposals");
    int rc = join_lanes(c); if (rc) return rc;
    NbList nb;
    for (int x = 0; x < 16; x++) nb.id[x] = x < n_proposals ? id_fB[x] : 0;
    k_draw_candidates<<<1, 32, 0, c->stream>>>(d_delta, d_full, likelihood_host, use_host_likelihood, n_proposals * GRAAL_N_CANDIDATES, GRAAL_N_CANDIDATES,
                                               u, nb, d_sel, d_sub_score, d_score);
    CHECK_LAUNCH(c);
    return 0;
}
int graal_draw_commit(graal_ctx* c, int base_slot, int first_cand_slot, int id_fA, const int32_t* id_fB, int n_proposals, int max_id,
                      const double* d_delta, const double* d_full, double likelihood_host, int use_host_likelihood, double u,
                      int32_t* d_sel, double* d_sub_score, double* d_score, const void* d_fetch_src, void* h_fetch_dst, size_t fetch_bytes) {
    NEED_STATE(c); NEED_SLOT(c, base_slot); NEED_SLOT(c, first_cand_slot); NEED_SLOT(c, first_cand_slot + GRAAL_N_CANDIDATES - 1);
    const int n = c->n_new;
    if (id_fA < 0 || id_fA >= n) return set_err(-1, "bin id out of range");
    if (base_slot >= first_cand_slot && base_slot < first_cand_slot + GRAAL_N_CANDIDATES) return set_err(-1, "source slot inside the destination range");
    for (int x = 0; x < n_proposals && id_fB; x++) if (id_fB[x] < 0 || id_fB[x] >= n) return set_err(-1, "bin id out of range");
    int rc = graal_draw_candidates(c, d_delta, d_full, likelihood_host, use_host_likelihood, n_proposals, id_fB, u, d_sel, d_sub_score, d_score); if (rc) return rc;
    const int keep = c->band_slot;
    cudaStream_t st = c->stream;
    if (d_fetch_src && h_fetch_dst && fetch_bytes) {
        // the step's results leave for the host right behind the draw, BEFORE the commit kernels: graal_fetch_wait returns while
        // the commit runs (contig ids of the relabelled slot are < n_contigs, + 3 for the candidate committed below)
        c->fetch_ncontigs = c->ncontigs_pending && c->h_ncontigs;
        void* h_dev = nullptr;
        const bool publish = c->publish_enable && (fetch_bytes % 8 == 0) && cudaHostGetDevicePointer(&h_dev, h_fetch_dst, 0) == cudaSuccess && h_dev == h_fetch_dst;
        if (!publish) (void)cudaGetLastError();
        if (publish) {
            c->fetch_seq++;
            k_publish<<<1, 256, 0, st>>>(static_cast<const double*>(d_fetch_src), static_cast<double*>(h_fetch_dst), (int)(fetch_bytes / 8), c->d_ints + 1,
                                         c->h_ncontigs, c->fetch_ncontigs ? 1 : 0, c->fetch_seq);
            CHECK_LAUNCH(c);
        } else {
            CUDA_OK(cudaMemcpyAsync(h_fetch_dst, d_fetch_src, fetch_bytes, cudaMemcpyDeviceToHost, st));
            if (c->fetch_ncontigs) CUDA_OK(cudaMemcpyAsync(c->h_ncontigs, c->d_ints + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
            CUDA_OK(cudaEventRecord(c->ev_fetch, st));
        }
        if (c->fetch_ncontigs) { c->fetch_mark = c->commits_total; c->ncontigs_pending = false; }
        c->fetch_pending = true; c->fetch_published = publish;
    }
    // test_copy_struct (cuda_lib_gl.py:1156-1183) of the candidate the device drew: rebuild it, copy it over the current slot
    k_build_candidates_sel<<<nblk(n, 128), 128, 0, st>>>(slot_ptr(c, base_slot), slot_ptr(c, first_cand_slot), slot_stride(c), c->ld, n, id_fA, d_sel, c->d_ints + 0, max_id);
    CHECK_LAUNCH(c);
    k_commit_sel<<<nblk(n, 128), 128, 0, st>>>(slot_ptr(c, first_cand_slot), slot_stride(c), slot_ptr(c, base_slot), c->ld, n, d_sel);
    CHECK_LAUNCH(c);
    for (int k = 0; k < GRAAL_N_CANDIDATES; k++) {
        const int sl = first_cand_slot + k;
        if (c->geo_base_slot == sl) c->geo_base_slot = -1;
        if (c->first_idx_slot == sl) c->first_idx_slot = -1;
        if (c->band_slot == sl) c->band_slot = -1;
        if (c->bound_slot == sl) { c->contig_bound = -1; c->ncontigs_pending = false; c->fetch_ncontigs = false; }
    }
    if (c->geo_base_slot == base_slot) c->geo_base_slot = -1;
    if (c->first_idx_slot == base_slot) c->first_idx_slot = -1;
    if (c->band_slot == base_slot) c->band_slot = -1;
    if (c->bound_slot == base_slot) { c->bound_commits++; c->commits_total++; if (c->contig_bound >= 0) c->contig_bound += 3; }
    if (keep == base_slot) {              // keep the cached band total in step with the committed candidate
        k_add_selected_sel<<<1, 1, 0, st>>>(c->d_scalars + 40, c->band_hist, d_sel); CHECK_LAUNCH(c);
        c->band_slot = base_slot;
    }
    return 0;
}
int graal_fetch_wait(graal_ctx* c) {
    if (!c) return set_err(-1, "null argument");
    if (!c->fetch_pending) return set_err(-1, "no fetch posted by graal_draw_commit");
    CUDA_OK(cudaSetDevice(c->device));
    if (c->fetch_published) {
        volatile int* seq = c->h_ncontigs + 8;
        for (long long spin = 0; *seq != c->fetch_seq; spin++) {
            if ((spin & 0xfffff) == 0xfffff) {                  // every ~1M polls: has the stream failed or drained without publishing?
                const cudaError_t q = cudaStreamQuery(c->stream);
                if (q != cudaSuccess && q != cudaErrorNotReady) return set_err(-2, "stream failed while waiting for the step's results: %s", cudaGetErrorString(q));
                if (q == cudaSuccess && *seq != c->fetch_seq) return set_err(-2, "the step's results were never published");
            }
        }
        __sync_synchronize();
    } else CUDA_OK(cudaEventSynchronize(c->ev_fetch));
    c->fetch_pending = false;
    // contig ids were < n_contigs when the readback left; every candidate committed since adds at most 3 (relabels only lower them)
    if (c->fetch_ncontigs) c->contig_bound = *c->h_ncontigs + 3 * (int)(c->commits_total - c->fetch_mark);
    c->fetch_ncontigs = false;
    return 0;
}
static double np_pairwise_sum(const double* a, long n) {
    if (n < 8) { double r = 0.0; for (long i = 0; i < n; i++) r += a[i]; return r; }      // (NumPy starts from a[0]: 0.0 + a[0] == a[0] except -0.0, never a weight)
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; j++) r[j] = a[j];
        long i;
        for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    }
    long n2 = n / 2; n2 -= n2 % 8;
    return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

// score[n] (n = 13 x neighbours) -> the indices id_ok[n_ok] that may be drawn and their normalised cumulative weights
// cdf[n_ok] (temperature 1: the power step of the reference is the identity); returns n_ok, *id_max = argmax(score).
// Returns -1 when a weight is not finite (the caller then runs the NumPy statements, which raise like the reference).
int graal_candidate_weights(const double* score, int n, int n_tmp, double* work, int32_t* id_ok, double* cdf, int32_t* id_max) {
    if (!score || !work || !id_ok || !cdf || !id_max || n <= 0 || n_tmp <= 0) return -2;
    int im = 0; double mn = score[0];
    bool nan = score[0] != score[0];
    for (int i = 1; i < n; i++) {
        if (score[i] != score[i]) nan = true;
        if (score[i] > score[im]) im = i;
        if (score[i] < mn) mn = score[i];
    }
    if (nan) return -1;
    *id_max = im;
    for (int i = 0; i < n; i++) work[i] = score[i] - mn;
    for (int i = n_tmp; i < n; i += n_tmp) { work[i] = 0.0; if (i + 1 < n) work[i + 1] = 0.0; }     // eject / flip of the later neighbours: duplicates of the first one's
    double mx = work[0];
    for (int i = 1; i < n; i++) if (work[i] > mx) mx = work[i];
    const double shift = mx - 30.0;                                                                 // thresh_overflow
    int n_ok = 0;
    for (int i = 0; i < n; i++) {
        double f = work[i] - shift;
        if (f < 0.0) f = 0.0;
        if (f > 0.0) { id_ok[n_ok] = i; cdf[n_ok] = f; n_ok++; }
    }
    if (n_ok == 0) return 0;
    double sum = np_pairwise_sum(cdf, n_ok);
    for (int i = 0; i < n_ok; i++) cdf[i] = cdf[i] / sum;
    sum = np_pairwise_sum(cdf, n_ok);                                                               // (power 1 / F_t with F_t == 1: identity) second normalisation
    for (int i = 0; i < n_ok; i++) cdf[i] = cdf[i] / sum;
    for (int i = 0; i < n_ok; i++) work[i] = cdf[i];                                                // sub_score (kept for the caller)
    double run = 0.0;
    for (int i = 0; i < n_ok; i++) { run += cdf[i]; cdf[i] = run; }
    const double last = cdf[n_ok - 1];
    if (!(last - last == 0.0)) return -1;
    for (int i = 0; i < n_ok; i++) cdf[i] = cdf[i] / last;
    return n_ok;
}

int graal_dist_histogram(graal_ctx* c, const int32_t* sub_id_c, const int32_t* sub_start_bp, const int32_t* sub_len_bp,
                         const int32_t* sub_pos, double max_dist_kb, double bin_kb, int n_bins, double* d_sum, int64_t* d_cnt) {
    if (!c || !c->rowptr) return set_err(-1, "level not bound");
    if (!sub_id_c || !sub_start_bp || !sub_len_bp || !sub_pos || !d_sum || !d_cnt || n_bins <= 0) return set_err(-1, "bad argument");
    CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    CUDA_OK(cudaMemsetAsync(d_sum, 0, (size_t)n_bins * sizeof(double), st));
    CUDA_OK(cudaMemsetAsync(d_cnt, 0, (size_t)n_bins * sizeof(int64_t), st));
    k_dist_hist<<<c->n_sm * 8, 256, 0, st>>>(c->rowptr, c->contacts, c->W, sub_id_c, sub_start_bp, sub_len_bp, sub_pos,
                                            max_dist_kb, bin_kb, n_bins, d_sum, reinterpret_cast<unsigned long long*>(d_cnt));
    CHECK_LAUNCH(c);
    return 0;
}

}  // extern "C"
