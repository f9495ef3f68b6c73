/*
 * lpg_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's
 * Local-Planar-Guidance hot path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (bts-fully-tf_b200/) never does and has no CPU fallback.
 *
 * Parity status: the reference (clarencechen/bts-fully-tf) ships no tests, golden
 * vectors or fixtures, and its arithmetic lives in TensorFlow, which cannot be
 * installed here.  This oracle is therefore pinned against the reference's OWN
 * SOURCE executed over a torch-CPU stand-in for the tf ops it calls
 * (oracle/tf_shim + tests/golden/make_golden.py -> tests/golden/ npz files).  TensorFlow's
 * own kernels' last-ulp rounding remains unpinned ("parity unpinned" w.r.t. a real
 * TF run).
 *
 * Layout: everything is contiguous NHWC like the reference (Keras channels_last):
 *   coef   (B, h, w, 3)   channel order [phi_raw, theta_raw, dist]  custom_layers.py:49
 *   out    (B, H, W, 1)   H = h*r, W = w*r                             custom_layers.py:32
 *   out_ds (B, H/d, W/d, 1) = out[:, ::d, ::d]                        bts_decoder.py:81,88
 *   feat   (B, h, w, C), kernel (1,1,C,3) HWIO == [C][3]              bts_decoder.py:79,86,93
 *
 * Build: oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#define ORACLE_API __attribute__((visibility("default")))

/* the python float `pi` becomes a float32 constant when it meets a float32 tensor
 * (custom_layers.py:49: inputs * 2 * pi ; inputs * pi / 3) */
static const float PI_F = 3.14159265358979323846f;
/* K.epsilon() == 1e-7, added as a float32 scalar (custom_layers.py:55) */
static const float EPS_F = 1e-7f;

/* ------------------------------------------------------------------------------------------
 * LocalPlanarGuidance.build  (custom_layers.py:30-45), literal, float32.
 *   v,u = meshgrid(linspace(0,W-1,W), linspace(0,H-1,H))   -> u = row index, v = column index
 *   v = (v % r - (r-1)/2) / r ; u likewise
 *   pixel_dir_unit = l2_normalize(stack([u, v, 1], -1), axis=3)
 * dir: (H, W, 3) float32.
 * ------------------------------------------------------------------------------------------ */
ORACLE_API void oracle_lpg_pixel_dir_f32(int H, int W, int r, float *dir)
{
    const float half = (float)((r - 1) / 2.0);
    const float rf = (float)r;
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            float u = (fmodf((float)y, rf) - half) / rf;
            float v = (fmodf((float)x, rf) - half) / rf;
            float one = 1.0f;
            float ss = (u * u + v * v) + one * one;   /* reduce_sum(square(x), axis=3) */
            float m = ss > 1e-12f ? ss : 1e-12f;       /* maximum(square_sum, epsilon)  */
            float inv = 1.0f / sqrtf(m);               /* rsqrt                         */
            float *d = dir + ((size_t)y * W + x) * 3;
            d[0] = u * inv; d[1] = v * inv; d[2] = one * inv;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * LocalPlanarGuidance.call (custom_layers.py:47-56), literal op order, float32, no FMA
 * contraction (this file is compiled with -ffp-contract=off).
 * ------------------------------------------------------------------------------------------ */
ORACLE_API void oracle_lpg_forward_f32(const float *coef, int B, int h, int w, int r, float *out)
{
    const int H = h * r, W = w * r;
    float *dir = (float *)malloc((size_t)H * W * 3 * sizeof(float));
    oracle_lpg_pixel_dir_f32(H, W, r, dir);
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int i = 0; i < h; ++i) {
            for (int j = 0; j < w; ++j) {
                const float *c = coef + (((size_t)b * h + i) * w + j) * 3;
                float phi = (c[0] * 2.0f) * PI_F;              /* :49 */
                float theta = (c[1] * PI_F) / 3.0f;            /* :49 */
                float n4 = c[2];
                float st = sinf(theta), ct = cosf(theta);
                float sp = sinf(phi), cp = cosf(phi);
                float n1 = st * cp, n2 = st * sp, n3 = ct;     /* :50 */
                for (int p = 0; p < r; ++p) {                  /* repeat_elements axis=1, :52 */
                    for (int q = 0; q < r; ++q) {              /* repeat_elements axis=2, :53 */
                        int y = i * r + p, x = j * r + q;
                        const float *d = dir + ((size_t)y * W + x) * 3;
                        float den = ((d[0] * n1 + d[1] * n2) + d[2] * n3) + EPS_F; /* :55 */
                        out[((size_t)b * H + y) * W + x] = n4 / den;               /* :56 */
                    }
                }
            }
        }
    }
    free(dir);
}

/* ------------------------------------------------------------------------------------------
 * Closed form in float64: exact arithmetic applied to the reference's float32 program, i.e.
 * the float32 CONSTANTS of the program are kept (fl32(pi), fl32(1e-7)) but every operation is
 * carried out in double and the direction vectors are exact.  This is the yard-stick the GPU
 * results are compared against (tolerance 1e-5 relative, see tests/).
 *   out = n4 / ( (a*n1 + b*n2 + n3)/sqrt(a^2+b^2+1) + eps ),
 *   a = ((y mod r) - (r-1)/2)/r  (rows), b likewise for columns.        SURVEY 8(a) a2-a5
 * coef may be given as float32 (coef32 != NULL) or float64 (coef64 != NULL).
 * den_out (nullable) receives the denominator so tests can apply the denominator-aware rule.
 * ------------------------------------------------------------------------------------------ */
static inline void decode_f64(double x0, double x1, double *sp, double *cp, double *st, double *ct)
{
    double phi = x0 * 2.0 * (double)PI_F;
    double theta = x1 * (double)PI_F / 3.0;
    *sp = sin(phi); *cp = cos(phi); *st = sin(theta); *ct = cos(theta);
}

static inline void dir_f64(int p, int q, int r, double *du, double *dv, double *dw)
{
    double a = ((double)p - (r - 1) / 2.0) / r;
    double b = ((double)q - (r - 1) / 2.0) / r;
    double inv = 1.0 / sqrt(a * a + b * b + 1.0);
    *du = a * inv; *dv = b * inv; *dw = inv;
}

ORACLE_API void oracle_lpg_forward_f64(const float *coef32, const double *coef64,
                                       int B, int h, int w, int r, double *out, double *den_out)
{
    const int H = h * r, W = w * r;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int i = 0; i < h; ++i) {
            for (int j = 0; j < w; ++j) {
                size_t ci = (((size_t)b * h + i) * w + j) * 3;
                double x0 = coef32 ? (double)coef32[ci] : coef64[ci];
                double x1 = coef32 ? (double)coef32[ci + 1] : coef64[ci + 1];
                double n4 = coef32 ? (double)coef32[ci + 2] : coef64[ci + 2];
                double sp, cp, st, ct;
                decode_f64(x0, x1, &sp, &cp, &st, &ct);
                double n1 = st * cp, n2 = st * sp, n3 = ct;
                for (int p = 0; p < r; ++p)
                    for (int q = 0; q < r; ++q) {
                        double du, dv, dw;
                        dir_f64(p, q, r, &du, &dv, &dw);
                        double den = du * n1 + dv * n2 + dw * n3 + (double)EPS_F;
                        size_t oi = ((size_t)b * H + (i * r + p)) * W + (j * r + q);
                        out[oi] = n4 / den;
                        if (den_out) den_out[oi] = den;
                    }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Down-sampled copy (bts_decoder.py:81,88): ds = full[:, ::d, ::d].  Generic over element size.
 * ------------------------------------------------------------------------------------------ */
ORACLE_API void oracle_downsample_f64(const double *full, int B, int H, int W, int d, double *ds)
{
    const int Hd = (H + d - 1) / d, Wd = (W + d - 1) / d;
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < Hd; ++y)
            for (int x = 0; x < Wd; ++x)
                ds[((size_t)b * Hd + y) * Wd + x] = full[((size_t)b * H + y * d) * W + x * d];
}

ORACLE_API void oracle_downsample_f32(const float *full, int B, int H, int W, int d, float *ds)
{
    const int Hd = (H + d - 1) / d, Wd = (W + d - 1) / d;
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < Hd; ++y)
            for (int x = 0; x < Wd; ++x)
                ds[((size_t)b * Hd + y) * Wd + x] = full[((size_t)b * H + y * d) * W + x * d];
}

/* ------------------------------------------------------------------------------------------
 * LPG backward, closed form, float64 (what TF autodiff of custom_layers.py:49-56 computes;
 * SURVEY 8(a) a6, a9).  g_full (B,H,W); g_ds nullable (B,H/d,W/d) is the gradient flowing
 * into the strided slice out[:, ::d, ::d] and is scattered back onto those pixels.
 *   G   = g_full + scatter(g_ds)
 *   g4  = sum_P G/den ; t = -G*n4/den^2 ; g1 = sum t*du ; g2 = sum t*dv ; g3 = sum t*dw
 *   gth = g1 ct cp + g2 ct sp - g3 st ;  gph = -g1 st sp + g2 st cp
 *   d/dx = [ 2*pi*gph , (pi/3)*gth , g4 ]
 * ------------------------------------------------------------------------------------------ */
ORACLE_API void oracle_lpg_backward_f64(const float *coef32, const double *coef64,
                                        const double *g_full, const double *g_ds, int d,
                                        int B, int h, int w, int r, double *g_coef)
{
    const int H = h * r, W = w * r;
    const int Hd = g_ds ? H / d : 0, Wd = g_ds ? W / d : 0;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int i = 0; i < h; ++i) {
            for (int j = 0; j < w; ++j) {
                size_t ci = (((size_t)b * h + i) * w + j) * 3;
                double x0 = coef32 ? (double)coef32[ci] : coef64[ci];
                double x1 = coef32 ? (double)coef32[ci + 1] : coef64[ci + 1];
                double n4 = coef32 ? (double)coef32[ci + 2] : coef64[ci + 2];
                double sp, cp, st, ct;
                decode_f64(x0, x1, &sp, &cp, &st, &ct);
                double n1 = st * cp, n2 = st * sp, n3 = ct;
                double g1 = 0, g2 = 0, g3 = 0, g4 = 0;
                for (int p = 0; p < r; ++p)
                    for (int q = 0; q < r; ++q) {
                        int y = i * r + p, x = j * r + q;
                        double G = g_full ? g_full[((size_t)b * H + y) * W + x] : 0.0;
                        if (g_ds && (y % d) == 0 && (x % d) == 0)
                            G += g_ds[((size_t)b * Hd + y / d) * Wd + x / d];
                        double du, dv, dw;
                        dir_f64(p, q, r, &du, &dv, &dw);
                        double den = du * n1 + dv * n2 + dw * n3 + (double)EPS_F;
                        double inv = 1.0 / den;
                        g4 += G * inv;
                        double t = -G * n4 * inv * inv;
                        g1 += t * du; g2 += t * dv; g3 += t * dw;
                    }
                double gth = g1 * ct * cp + g2 * ct * sp - g3 * st;
                double gph = -g1 * st * sp + g2 * st * cp;
                g_coef[ci] = 2.0 * (double)PI_F * gph;
                g_coef[ci + 1] = ((double)PI_F / 3.0) * gth;
                g_coef[ci + 2] = g4;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * reduction_{8x8,4x4,2x2} head (bts_decoder.py:79,86,93): Conv2D(3, 1x1, sigmoid, no bias).
 *   z[p,k] = sum_c feat[p,c] * W[c,k] ;  x = 1/(1+exp(-z))
 * npix = B*h*w.  Float64 math; feat/W may be given as float32 or float64.
 * ------------------------------------------------------------------------------------------ */
ORACLE_API void oracle_head_forward_f64(const float *feat32, const double *feat64,
                                        const float *w32, const double *w64,
                                        size_t npix, int C, double *coef)
{
#pragma omp parallel for schedule(static)
    for (size_t p = 0; p < npix; ++p) {
        double z[3] = {0, 0, 0};
        for (int c = 0; c < C; ++c) {
            double f = feat32 ? (double)feat32[p * C + c] : feat64[p * C + c];
            for (int k = 0; k < 3; ++k)
                z[k] += f * (w32 ? (double)w32[c * 3 + k] : w64[c * 3 + k]);
        }
        for (int k = 0; k < 3; ++k) coef[p * 3 + k] = 1.0 / (1.0 + exp(-z[k]));
    }
}

/* head backward (autodiff of the above):
 *   dz = g_coef * x*(1-x) ; g_w[c,k] = sum_p feat[p,c]*dz[p,k] ; g_feat[p,c] = sum_k dz[p,k]*W[c,k] */
ORACLE_API void oracle_head_backward_f64(const float *feat32, const double *feat64,
                                         const float *w32, const double *w64,
                                         const double *coef, const double *g_coef,
                                         size_t npix, int C, double *g_feat, double *g_w)
{
    for (int i = 0; i < C * 3; ++i) g_w[i] = 0.0;
    for (size_t p = 0; p < npix; ++p) {
        double dz[3];
        for (int k = 0; k < 3; ++k) {
            double x = coef[p * 3 + k];
            dz[k] = g_coef[p * 3 + k] * x * (1.0 - x);
        }
        for (int c = 0; c < C; ++c) {
            double f = feat32 ? (double)feat32[p * C + c] : feat64[p * C + c];
            double acc = 0.0;
            for (int k = 0; k < 3; ++k) {
                g_w[c * 3 + k] += f * dz[k];
                acc += dz[k] * (w32 ? (double)w32[c * 3 + k] : w64[c * 3 + k]);
            }
            if (g_feat) g_feat[p * C + c] = acc;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Closed-form float32 forward + backward with FMA contraction disabled: the scalar "port"
 * timed as a single-core CPU figure next to the literal torch restatement (bench.py).
 * ------------------------------------------------------------------------------------------ */
ORACLE_API void oracle_lpg_fwdbwd_f32(const float *coef, const float *g_full,
                                      int B, int h, int w, int r, float *out, float *g_coef)
{
    const int H = h * r, W = w * r;
    float du[64], dv[64], dw[64];
    for (int p = 0; p < r; ++p)
        for (int q = 0; q < r; ++q) {
            float a = ((float)p - (float)((r - 1) / 2.0)) / (float)r;
            float b = ((float)q - (float)((r - 1) / 2.0)) / (float)r;
            float inv = 1.0f / sqrtf(a * a + b * b + 1.0f);
            du[p * r + q] = a * inv; dv[p * r + q] = b * inv; dw[p * r + q] = inv;
        }
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int i = 0; i < h; ++i) {
            for (int j = 0; j < w; ++j) {
                size_t ci = (((size_t)b * h + i) * w + j) * 3;
                float phi = (coef[ci] * 2.0f) * PI_F, theta = (coef[ci + 1] * PI_F) / 3.0f, n4 = coef[ci + 2];
                float st = sinf(theta), ct = cosf(theta), sp = sinf(phi), cp = cosf(phi);
                float n1 = st * cp, n2 = st * sp, n3 = ct;
                float g1 = 0, g2 = 0, g3 = 0, g4 = 0;
                for (int p = 0; p < r; ++p)
                    for (int q = 0; q < r; ++q) {
                        size_t oi = ((size_t)b * H + (i * r + p)) * W + (j * r + q);
                        int k = p * r + q;
                        float den = ((du[k] * n1 + dv[k] * n2) + dw[k] * n3) + EPS_F;
                        out[oi] = n4 / den;
                        if (g_full) {
                            float inv = 1.0f / den, G = g_full[oi];
                            g4 += G * inv;
                            float t = -G * n4 * inv * inv;
                            g1 += t * du[k]; g2 += t * dv[k]; g3 += t * dw[k];
                        }
                    }
                if (g_full) {
                    g_coef[ci] = 2.0f * PI_F * (-g1 * st * sp + g2 * st * cp);
                    g_coef[ci + 1] = (PI_F / 3.0f) * (g1 * ct * cp + g2 * ct * sp - g3 * st);
                    g_coef[ci + 2] = g4;
                }
            }
        }
    }
}

ORACLE_API int oracle_version(void) { return 1; }
