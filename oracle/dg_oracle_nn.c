/*
 * dg_oracle_nn.c -- CPU restatement of dream-go's `dg_nn::forward`.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (dream_go_b200/,
 * include/) may link, load or call this file; only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() use it, and
 * there only as the checker.
 *
 * What it restates (paths relative to the reference tree):
 *   src/libdg_nn/graph.rs:123-158              forward(): up -> residual* -> {value, policy}
 *   src/libdg_nn/layers/up_block.rs:48-57      relu(conv3x3(features) + b)
 *   src/libdg_nn/layers/residual_block.rs:40-78
 *        y = relu(conv1(x)+b1); out = relu(g*conv2(y) + (1-g)*x + fp16(g*b2))
 *   src/libdg_nn/layers/conv2d.rs:99-142,171-220   NHWC activations, KRSC filters,
 *        cross-correlation, pad 1, stride 1, y = act(a1*conv + a2*z + b)
 *   src/libdg_nn/layers/dense.rs:79-152,173-220    FC weight stored [in][out]
 *   src/libdg_nn/layers/policy_head.rs:43-103  softmax(tau*(W.flat(relu(conv))) + fp16(tau*b))
 *   src/libdg_nn/layers/value_head.rs:40-84    tanh(w.flat(relu(conv)) + b)
 *   src/libdg_utils/b85.rs:17-170              RFC 1924 base85 tensor coding
 *   src/libdg_utils/types/fp16.rs:50-75        IEEE binary16 <-> binary32 (RNE)
 *
 * The arithmetic of the reference lives in NVIDIA cuDNN (closed source, not in
 * Cargo.lock; README.md:11 says "cuDNN v8 or higher").  This file restates
 * cuDNN's documented semantics of cudnnConvolutionBiasActivationForward,
 * cudnnSoftmaxForward(ACCURATE, INSTANCE), cudnnActivationForward(TANH) and
 * cudnnScaleTensor.  Every tensor the reference materialises in fp16 is
 * rounded to fp16 here at the same point; accumulation is done in double so
 * that this oracle is the "ideal" fp16-storage network.
 *
 * Pinning: the reference's own known-answer tests for this path
 * (conv2d.rs:253-291, dense.rs:244-300, loader.rs:124-142, b85.rs:168-220,
 * fp16.rs:77-92) are replayed against this file in tests/test_oracle_kat.py.
 * The reference holds NO whole-network golden vector and ships no weights, so
 * whole-network parity is pinned only through those layer-level KATs
 * ("parity unpinned" at network level -- see DESIGN.md).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define DG_BOARD 19
#define DG_POINTS 361
#define DG_POLICY 362

/* ------------------------------------------------------------------ fp16 */

/* binary32 -> binary16, round-to-nearest-even (llvm.convert.to.fp16.f32,
 * src/libdg_utils/types/fp16.rs:50-56). */
uint16_t dg_oracle_f32_to_f16(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t absx = x & 0x7fffffffu;
    if (absx >= 0x7f800000u) { /* inf / nan */
        return (uint16_t)(sign | 0x7c00u | ((absx > 0x7f800000u) ? (0x200u | ((absx >> 13) & 0x3ffu)) : 0));
    }
    if (absx >= 0x477ff000u) { /* rounds to >= 65520 -> inf */
        return (uint16_t)(sign | 0x7c00u);
    }
    if (absx < 0x33000001u) { /* < 2^-25 (or == 2^-25, tie to even -> 0) */
        return (uint16_t)sign;
    }
    int32_t exp = (int32_t)(absx >> 23) - 127;
    uint32_t man = (absx & 0x7fffffu) | 0x800000u;
    uint32_t shift, half_bits;
    if (exp < -14) { /* subnormal half */
        shift = (uint32_t)(13 + (-14 - exp));
        half_bits = 0;
    } else {
        shift = 13;
        half_bits = (uint32_t)(exp + 15) << 10;
        man &= 0x7fffffu;
    }
    uint32_t q = man >> shift;
    uint32_t rem = man & ((1u << shift) - 1u);
    uint32_t halfway = 1u << (shift - 1);
    uint32_t out = half_bits + q;
    if (rem > halfway || (rem == halfway && (out & 1u))) out += 1; /* carries into exponent correctly */
    return (uint16_t)(sign | out);
}

float dg_oracle_f16_to_f32(uint16_t h) {
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t man = h & 0x3ffu;
    uint32_t x;
    if (exp == 0) {
        if (man == 0) {
            x = sign;
        } else { /* subnormal */
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            x = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) {
        x = sign | 0x7f800000u | (man << 13);
    } else {
        x = sign | ((exp + 112u) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &x, 4);
    return f;
}

static inline float round_f16(double v) {
    return dg_oracle_f16_to_f32(dg_oracle_f32_to_f16((float)v));
}

/* ------------------------------------------------------------------ base85 */

static const char B85_ALPHABET[86] =
    "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz!#$%&()*+-;<=>?@^_`{|}~";

/* Decodes groups of five characters into big-endian 32-bit words and emits the
 * four bytes of each word in stream order (src/libdg_utils/b85.rs:101-139; the
 * per-type `FromB85` impls at :46-94 all amount to "little-endian elements in
 * stream order").  A trailing partial group is dropped, as the reference does
 * (`break 'outer`).  Returns the number of bytes written, or -1 on an invalid
 * character. */
long dg_oracle_b85_decode(const char* in, long n_in, uint8_t* out, long cap) {
    int8_t table[256];
    memset(table, -1, sizeof table);
    for (int i = 0; i < 85; i++) table[(uint8_t)B85_ALPHABET[i]] = (int8_t)i;
    long n_out = 0;
    for (long i = 0; i + 5 <= n_in; i += 5) {
        uint32_t acc = 0;
        for (int j = 0; j < 5; j++) {
            int8_t d = table[(uint8_t)in[i + j]];
            if (d < 0) return -1;
            acc = 85u * acc + (uint32_t)d;
        }
        if (n_out + 4 > cap) return -2;
        out[n_out++] = (uint8_t)(acc >> 24);
        out[n_out++] = (uint8_t)(acc >> 16);
        out[n_out++] = (uint8_t)(acc >> 8);
        out[n_out++] = (uint8_t)acc;
    }
    /* an invalid character inside a trailing partial group is still an error */
    for (long i = n_in - (n_in % 5); i < n_in; i++)
        if (table[(uint8_t)in[i]] < 0) return -1;
    return n_out;
}

/* Encodes raw bytes (length a multiple of 4) -- the inverse of the above and
 * what contrib/trainer/dream_tf/hooks/dump.py:54-61 produces. */
long dg_oracle_b85_encode(const uint8_t* in, long n_in, char* out, long cap) {
    if (n_in % 4) return -1;
    long n_out = 0;
    for (long i = 0; i < n_in; i += 4) {
        uint32_t acc = ((uint32_t)in[i] << 24) | ((uint32_t)in[i + 1] << 16) | ((uint32_t)in[i + 2] << 8) | in[i + 3];
        if (n_out + 5 > cap) return -2;
        char tmp[5];
        for (int j = 4; j >= 0; j--) { tmp[j] = B85_ALPHABET[acc % 85u]; acc /= 85u; }
        memcpy(out + n_out, tmp, 5);
        n_out += 5;
    }
    return n_out;
}

/* ------------------------------------------------------------------ layers */

/* y[n,p,k] = act(a1 * sum_{r,s,c} x[n, p+(r-1,s-1), c] * w[k,r,s,c] + a2 * z[n,p,k] + b[k]),
 * rounded to fp16.  Restates cudnnConvolutionBiasActivationForward as driven by
 * src/libdg_nn/layers/conv2d.rs:171-220 (CrossCorrelation, pad = size/2,
 * stride 1, NHWC x/y/z, KRSC w).  `x`, `z`, `y` hold fp16-representable floats.
 * wh = board edge (19 in the network, 3 in the reference KAT). */
void dg_oracle_conv3x3(const float* x, int n, int wh, int cin,
                       const uint16_t* w_krsc, const uint16_t* bias, int cout,
                       float a1, float a2, const float* z, int relu, float* y) {
    /* transpose weights to [tap][c][k] floats so the inner loop runs over k */
    float* wt = (float*)malloc(sizeof(float) * 9u * (size_t)cin * (size_t)cout);
    float* bf = (float*)malloc(sizeof(float) * (size_t)cout);
    for (int k = 0; k < cout; k++) {
        bf[k] = dg_oracle_f16_to_f32(bias[k]);
        for (int t = 0; t < 9; t++)
            for (int c = 0; c < cin; c++)
                wt[((size_t)t * cin + c) * cout + k] = dg_oracle_f16_to_f32(w_krsc[((size_t)k * 9 + t) * cin + c]);
    }
    const int pts = wh * wh;
#pragma omp parallel
    {
        double* acc = (double*)malloc(sizeof(double) * (size_t)cout);
#pragma omp for schedule(static)
        for (long np = 0; np < (long)n * pts; np++) {
            const int b = (int)(np / pts), p = (int)(np % pts);
            const int py = p / wh, px = p % wh;
            for (int k = 0; k < cout; k++) acc[k] = 0.0;
            for (int r = 0; r < 3; r++) {
                const int yy = py + r - 1;
                if (yy < 0 || yy >= wh) continue;
                for (int s = 0; s < 3; s++) {
                    const int xx = px + s - 1;
                    if (xx < 0 || xx >= wh) continue;
                    const float* xin = x + ((size_t)b * pts + (size_t)yy * wh + xx) * cin;
                    const float* wtap = wt + (size_t)(r * 3 + s) * cin * cout;
                    for (int c = 0; c < cin; c++) {
                        const double xv = xin[c];
                        if (xv == 0.0) continue;
                        const float* wrow = wtap + (size_t)c * cout;
                        for (int k = 0; k < cout; k++) acc[k] += xv * (double)wrow[k];
                    }
                }
            }
            float* yo = y + (size_t)np * cout;
            const float* zo = z ? z + (size_t)np * cout : NULL;
            for (int k = 0; k < cout; k++) {
                double v = (double)a1 * acc[k] + (double)bf[k];
                if (zo) v += (double)a2 * (double)zo[k];
                if (relu && !(v > 0.0)) v = 0.0; /* NaN-non-propagating relu, activation_descriptor.rs:115-121 */
                yo[k] = round_f16(v);
            }
        }
        free(acc);
    }
    free(wt);
    free(bf);
}

/* y[n,o] = act(a1 * sum_i x[n,i] * w[i][o] + b[o]) rounded to fp16; identity
 * activation unless relu != 0.  Restates the 1x1-convolution trick of
 * src/libdg_nn/layers/dense.rs:137-152,197-220 (weights stored [in][out] in the
 * file and transposed on first use, :79-96,173-195). */
void dg_oracle_dense(const float* x, int n, int n_in, const uint16_t* w_in_out, const uint16_t* bias,
                     int n_out, float a1, int relu, float* y) {
#pragma omp parallel for schedule(static)
    for (int b = 0; b < n; b++) {
        double* acc = (double*)calloc((size_t)n_out, sizeof(double));
        for (int i = 0; i < n_in; i++) {
            const double xv = x[(size_t)b * n_in + i];
            if (xv == 0.0) continue;
            const uint16_t* wrow = w_in_out + (size_t)i * n_out;
            for (int o = 0; o < n_out; o++) acc[o] += xv * (double)dg_oracle_f16_to_f32(wrow[o]);
        }
        for (int o = 0; o < n_out; o++) {
            double v = (double)a1 * acc[o] + (double)dg_oracle_f16_to_f32(bias[o]);
            if (relu && !(v > 0.0)) v = 0.0;
            y[(size_t)b * n_out + o] = round_f16(v);
        }
        free(acc);
    }
}

/* ------------------------------------------------------------------ network */

typedef struct {
    int32_t num_blocks;     /* residual blocks found in the file (graph.rs:76-96) */
    int32_t channels;       /* num_channels:0, default 128 (layers/common.rs:22) */
    int32_t features;       /* 32 input planes (libdg_go/utils/features.rs:88-90) */
    int32_t policy_samples; /* num_samples:0, default 8 (layers/common.rs:25) */
    int32_t value_samples;  /* 2 (value_head.rs:42) */
    float tau;              /* 1 / SOFTMAX_TEMPERATURE (policy_head.rs:46) */
    const uint16_t* up_w;   /* 01_upsample/conv_1:0          [C][3][3][F] */
    const uint16_t* up_b;   /* 01_upsample/conv_1/offset:0   [C] */
    const uint16_t* const* res_w1; /* NN_residual/conv_1:0          [C][3][3][C], per block */
    const uint16_t* const* res_b1;
    const uint16_t* const* res_w2;
    const uint16_t* const* res_b2;
    const float* res_gate;  /* NN_residual/alpha:0, default 0.5 */
    const uint16_t* pol_conv_w; /* NNp_policy/conv_1:0    [S][3][3][C] */
    const uint16_t* pol_conv_b;
    const uint16_t* pol_fc_w;   /* NNp_policy/linear_1:0  [361*S][362] */
    const uint16_t* pol_fc_b;
    const uint16_t* val_conv_w; /* NNv_value/conv_1:0     [2][3][3][C] */
    const uint16_t* val_conv_b;
    const uint16_t* val_fc_w;   /* NNv_value/linear_2:0   [722][1] */
    const uint16_t* val_fc_b;
} dg_oracle_net;

static uint16_t* scale_bias_f16(const uint16_t* b, int n, float a) {
    /* cudnnScaleTensor on the fp16 offset, in place, once (residual_block.rs:72-74,
     * policy_head.rs:87-89): the scaled bias is itself rounded to fp16. */
    uint16_t* out = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)n);
    for (int i = 0; i < n; i++) out[i] = dg_oracle_f32_to_f16((float)((double)a * (double)dg_oracle_f16_to_f32(b[i])));
    return out;
}

/* features: [batch][361][32] fp16 bits (NHWC).  value_out: [batch] fp16 bits,
 * policy_out: [batch][362] fp16 bits.  tower_out (optional): final tower
 * activation [batch][361][C] as fp16 bits; block_out (optional): activation
 * after the up layer and after every block, [(num_blocks+1)][batch][361][C]. */
int dg_oracle_forward(const dg_oracle_net* net, const uint16_t* features, int batch,
                      uint16_t* value_out, uint16_t* policy_out, uint16_t* tower_out, uint16_t* block_out) {
    const int C = net->channels, F = net->features, S = net->policy_samples, V = net->value_samples;
    const size_t act = (size_t)batch * DG_POINTS * (size_t)C;
    float* x0 = (float*)malloc(sizeof(float) * (size_t)batch * DG_POINTS * (size_t)F);
    float* a = (float*)malloc(sizeof(float) * act);
    float* y = (float*)malloc(sizeof(float) * act);
    float* a2 = (float*)malloc(sizeof(float) * act);
    if (!x0 || !a || !y || !a2) return -1;
    for (size_t i = 0; i < (size_t)batch * DG_POINTS * (size_t)F; i++) x0[i] = dg_oracle_f16_to_f32(features[i]);

    dg_oracle_conv3x3(x0, batch, DG_BOARD, F, net->up_w, net->up_b, C, 1.0f, 0.0f, NULL, 1, a);
    if (block_out)
        for (size_t i = 0; i < act; i++) block_out[i] = dg_oracle_f32_to_f16(a[i]);

    for (int i = 0; i < net->num_blocks; i++) {
        const float g = net->res_gate[i];
        uint16_t* b2s = scale_bias_f16(net->res_b2[i], C, g);
        dg_oracle_conv3x3(a, batch, DG_BOARD, C, net->res_w1[i], net->res_b1[i], C, 1.0f, 0.0f, NULL, 1, y);
        dg_oracle_conv3x3(y, batch, DG_BOARD, C, net->res_w2[i], b2s, C, g, 1.0f - g, a, 1, a2);
        free(b2s);
        float* t = a; a = a2; a2 = t;
        if (block_out)
            for (size_t j = 0; j < act; j++) block_out[(size_t)(i + 1) * act + j] = dg_oracle_f32_to_f16(a[j]);
    }
    if (tower_out)
        for (size_t i = 0; i < act; i++) tower_out[i] = dg_oracle_f32_to_f16(a[i]);

    /* policy head */
    float* p1 = (float*)malloc(sizeof(float) * (size_t)batch * DG_POINTS * (size_t)S);
    float* p2 = (float*)malloc(sizeof(float) * (size_t)batch * DG_POLICY);
    dg_oracle_conv3x3(a, batch, DG_BOARD, C, net->pol_conv_w, net->pol_conv_b, S, 1.0f, 0.0f, NULL, 1, p1);
    uint16_t* pbs = scale_bias_f16(net->pol_fc_b, DG_POLICY, net->tau);
    dg_oracle_dense(p1, batch, DG_POINTS * S, net->pol_fc_w, pbs, DG_POLICY, net->tau, 0, p2);
    free(pbs);
    for (int b = 0; b < batch; b++) { /* cudnnSoftmaxForward ACCURATE / INSTANCE, softmax.rs:58-79 */
        const float* l = p2 + (size_t)b * DG_POLICY;
        double m = l[0], sum = 0.0;
        for (int o = 1; o < DG_POLICY; o++) if (l[o] > m) m = l[o];
        for (int o = 0; o < DG_POLICY; o++) sum += exp((double)l[o] - m);
        for (int o = 0; o < DG_POLICY; o++)
            policy_out[(size_t)b * DG_POLICY + o] = dg_oracle_f32_to_f16((float)(exp((double)l[o] - m) / sum));
    }
    free(p1);
    free(p2);

    /* value head */
    float* v1 = (float*)malloc(sizeof(float) * (size_t)batch * DG_POINTS * (size_t)V);
    float* v2 = (float*)malloc(sizeof(float) * (size_t)batch);
    dg_oracle_conv3x3(a, batch, DG_BOARD, C, net->val_conv_w, net->val_conv_b, V, 1.0f, 0.0f, NULL, 1, v1);
    dg_oracle_dense(v1, batch, DG_POINTS * V, net->val_fc_w, net->val_fc_b, 1, 1.0f, 0, v2);
    for (int b = 0; b < batch; b++) value_out[b] = dg_oracle_f32_to_f16((float)tanh((double)v2[b]));
    free(v1);
    free(v2);

    free(x0); free(a); free(y); free(a2);
    return 0;
}

int dg_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
