/* Plain-C restatement of the reference's HiFiGAN generator forward.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/hifigan_oracle.py).  Independent of
 * torch: direct convolutions with double accumulation, fp32 storage between
 * layers like the reference.  Follows, in /root/reference:
 *   src/iris/hifigan_pretrained.py:61-62    _get_padding
 *   src/iris/hifigan_pretrained.py:64-71    ResBlock.forward
 *   src/iris/hifigan_pretrained.py:123-143  HiFiGANModel.forward
 * and torch's definitions of Conv1d / ConvTranspose1d / weight_norm(dim=0)
 * (third-party; torch 2.9.1 pinned by the reference's uv.lock).
 *
 * Layout is the reference's: activations [B][C][L], Conv1d weight
 * [C_out][C_in][k], ConvTranspose1d weight [C_in][C_out][k].
 *
 * Build: gcc -O3 -fopenmp -shared -fPIC -o oracle/libhifigan_ref.so oracle/hifigan_ref.c -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define HFGREF_MAX 16

typedef struct {
    int in_channels;
    int upsample_initial_channel;
    int num_upsamples;
    int upsample_rates[HFGREF_MAX];
    int upsample_kernel_sizes[HFGREF_MAX];
    int num_kernels;
    int resblock_kernel_sizes[HFGREF_MAX];
    int num_dilations[HFGREF_MAX];
    int resblock_dilations[HFGREF_MAX][HFGREF_MAX];
} hfgref_config;

/* w = v * (g / ||v||), norm over all dims but 0 (rows). */
void hfgref_fold_weight_norm(const float* g, const float* v, int rows, int cols, float* w) {
    for (int r = 0; r < rows; ++r) {
        double s = 0.0;
        for (int c = 0; c < cols; ++c) s += (double)v[(size_t)r * cols + c] * v[(size_t)r * cols + c];
        float scale = g[r] / (float)sqrt(s);
        for (int c = 0; c < cols; ++c) w[(size_t)r * cols + c] = v[(size_t)r * cols + c] * scale;
    }
}

void hfgref_leaky_relu(const float* x, size_t n, float slope, float* y) {
    for (size_t i = 0; i < n; ++i) y[i] = x[i] > 0.f ? x[i] : x[i] * slope;
}

/* y[b][co][t] = bias[co] + sum_ci sum_j w[co][ci][j] * x[b][ci][t - pad + j*dil] (zero outside) */
void hfgref_conv1d(const float* x, int B, int Cin, int L, const float* w, const float* bias,
                   int Cout, int k, int dil, int pad, float* y) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int co = 0; co < Cout; ++co) {
            double* acc = (double*)malloc(sizeof(double) * (size_t)L);
            for (int t = 0; t < L; ++t) acc[t] = bias ? (double)bias[co] : 0.0;
            for (int ci = 0; ci < Cin; ++ci) {
                const float* xr = x + ((size_t)b * Cin + ci) * L;
                for (int j = 0; j < k; ++j) {
                    double wv = w[((size_t)co * Cin + ci) * k + j];
                    int off = j * dil - pad;
                    int t0 = off < 0 ? -off : 0;
                    int t1 = L - off < L ? L - off : L;
                    for (int t = t0; t < t1; ++t) acc[t] += wv * (double)xr[t + off];
                }
            }
            float* yr = y + ((size_t)b * Cout + co) * L;
            for (int t = 0; t < L; ++t) yr[t] = (float)acc[t];
            free(acc);
        }
}

/* y[b][co][t*stride - pad + j] += x[b][ci][t] * w[ci][co][j];  Lout = (L-1)*stride - 2*pad + k */
void hfgref_conv_transpose1d(const float* x, int B, int Cin, int L, const float* w, const float* bias,
                             int Cout, int k, int stride, int pad, float* y) {
    int Lout = (L - 1) * stride - 2 * pad + k;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int co = 0; co < Cout; ++co) {
            double* acc = (double*)malloc(sizeof(double) * (size_t)Lout);
            for (int t = 0; t < Lout; ++t) acc[t] = bias ? (double)bias[co] : 0.0;
            for (int ci = 0; ci < Cin; ++ci) {
                const float* xr = x + ((size_t)b * Cin + ci) * L;
                const float* wr = w + ((size_t)ci * Cout + co) * k;
                for (int t = 0; t < L; ++t) {
                    double xv = xr[t];
                    for (int j = 0; j < k; ++j) {
                        int o = t * stride - pad + j;
                        if (o >= 0 && o < Lout) acc[o] += xv * (double)wr[j];
                    }
                }
            }
            float* yr = y + ((size_t)b * Cout + co) * Lout;
            for (int t = 0; t < Lout; ++t) yr[t] = (float)acc[t];
            free(acc);
        }
}

static int get_padding(int k, int d) { return (k * d - d) / 2; }

/* weights: folded fp32 tensors, two pointers (weight, bias) per conv in the
 * order conv_pre, ups[0..], resblocks[n].{convs1[m], convs2[m]} for n, m
 * ascending (c1 then c2 per dilation), conv_post.
 * mel [B][in_channels][T] -> out [B][T*hop].  Returns 0, or -1 on bad args. */
int hfgref_forward(const hfgref_config* cfg, const float* const* weights, const float* mel,
                   int B, int T, float* out) {
    if (!cfg || !weights || !mel || !out || B <= 0 || T <= 0) return -1;
    int wi = 0;
    int C = cfg->upsample_initial_channel;
    size_t L = (size_t)T;
    float* x = (float*)malloc(sizeof(float) * (size_t)B * C * L);
    hfgref_conv1d(mel, B, cfg->in_channels, T, weights[0], weights[1], C, 7, 1, 3, x);
    wi = 2;
    int rb_base = 2 + 2 * cfg->num_upsamples;
    int rb_w = rb_base;
    for (int i = 0; i < cfg->num_upsamples; ++i) {
        int u = cfg->upsample_rates[i], k = cfg->upsample_kernel_sizes[i];
        int Cn = cfg->upsample_initial_channel >> (i + 1);
        size_t n_in = (size_t)B * C * L;
        hfgref_leaky_relu(x, n_in, 0.1f, x);
        size_t Ln = L * u;
        float* y = (float*)malloc(sizeof(float) * (size_t)B * Cn * Ln);
        hfgref_conv_transpose1d(x, B, C, (int)L, weights[wi], weights[wi + 1], Cn, k, u, (k - u) / 2, y);
        wi += 2;
        free(x);
        C = Cn; L = Ln;
        size_t n = (size_t)B * C * L;
        float* xs = (float*)calloc(n, sizeof(float));
        float* r = (float*)malloc(sizeof(float) * n);
        float* xt = (float*)malloc(sizeof(float) * n);
        float* xt2 = (float*)malloc(sizeof(float) * n);
        for (int j = 0; j < cfg->num_kernels; ++j) {
            int kk = cfg->resblock_kernel_sizes[j];
            memcpy(r, y, sizeof(float) * n);
            for (int m = 0; m < cfg->num_dilations[j]; ++m) {
                int d = cfg->resblock_dilations[j][m];
                hfgref_leaky_relu(r, n, 0.1f, xt);
                hfgref_conv1d(xt, B, C, (int)L, weights[rb_w], weights[rb_w + 1], C, kk, d, get_padding(kk, d), xt2);
                hfgref_leaky_relu(xt2, n, 0.1f, xt2);
                hfgref_conv1d(xt2, B, C, (int)L, weights[rb_w + 2], weights[rb_w + 3], C, kk, 1, get_padding(kk, 1), xt);
                rb_w += 4;
                for (size_t e = 0; e < n; ++e) r[e] = xt[e] + r[e];
            }
            if (j == 0) memcpy(xs, r, sizeof(float) * n);
            else for (size_t e = 0; e < n; ++e) xs[e] += r[e];
        }
        float nk = (float)cfg->num_kernels;
        for (size_t e = 0; e < n; ++e) xs[e] = xs[e] / nk;
        free(r); free(xt); free(xt2); free(y);
        x = xs;
    }
    size_t n = (size_t)B * C * L;
    hfgref_leaky_relu(x, n, 0.1f, x);
    hfgref_conv1d(x, B, C, (int)L, weights[rb_w], weights[rb_w + 1], 1, 7, 1, 3, out);
    for (size_t e = 0; e < (size_t)B * L; ++e) out[e] = tanhf(out[e]);
    free(x);
    return 0;
}
