// kernels_conv.cu — gold's Atari policy (pkg/v1/agent/deepq/policy.go:41-47,62-71): Conv2D 32x8x8/4 -> 64x4x4/2 ->
// 64x3x3/1 (pad 1, ReLU, no bias: goro layer/conv2d.go:61-160) -> Flatten -> FC 512 -> FC A, in fp32 on the CUDA
// cores.  SURVEY 8(f) rank 4: the reference marks this policy "not yet proven to converge"
// (experiments/pong/main.go:11-13); it is implemented for completeness of the drop-in surface, with the
// reference's arithmetic in the forward pass, and is NOT a tuned path (the tcgen05 kernels serve the MLP policies
// the headline metric is quoted on).
//
// Forward arithmetic = gorgonia's Conv2d (nn.go:236-314): im2col row (order c, ky, kx; zeros in the padding) dotted
// with the filter row by gonum's sgemmSerialNotTrans, i.e. f32.DotUnitary's lane order: two 4-lane accumulators
// over 16-element chunks, folded as (l0+l1)+(l2+l3) -- reproduced here so that Q matches the oracle bit for bit.
#include "common.cuh"

namespace gq {

namespace {

// y = Rectify(conv(x, f)): one thread per output element (b, o, oy, ox)
__global__ void __launch_bounds__(256)
conv_fwd_kernel(ConvShape c, int n, const float *__restrict__ x, const float *__restrict__ f, float *__restrict__ y,
                uint8_t *__restrict__ mask)
{
    const int S = c.OH * c.OW;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * c.O * S) return;
    const int s = (int)(t % S), o = (int)((t / S) % c.O), b = (int)(t / ((int64_t)S * c.O));
    const int oy = s / c.OW, ox = s - oy * c.OW;
    const float *im = x + (int64_t)b * c.C * c.H * c.W;
    const float *fr = f + (int64_t)o * c.C * c.KH * c.KW;
    const int K = c.C * c.KH * c.KW, KK = c.KH * c.KW;
    float sl[4] = {0.f, 0.f, 0.f, 0.f}, pl[4] = {0.f, 0.f, 0.f, 0.f};
    const int full = K & ~15;
    for (int k = 0; k < K; k++) {
        const int ch = k / KK, r = k - ch * KK, ky = r / c.KW, kx = r - ky * c.KW;
        const int iy = -c.pad + ky + oy * c.stride, ix = -c.pad + kx + ox * c.stride;
        const float xv = (iy < 0 || iy >= c.H || ix < 0 || ix >= c.W) ? 0.f : __ldg(im + ((int64_t)ch * c.H + iy) * c.W + ix);
        const float p = __fmul_rn(xv, __ldg(fr + k));
        if (k < full) {
            const int q = k & 3, grp = (k >> 2) & 3;
            if (grp & 1) pl[q] = __fadd_rn(pl[q], p);
            else sl[q] = __fadd_rn(sl[q], p);
        } else if (k < (K & ~3)) {
            sl[k & 3] = __fadd_rn(sl[k & 3], p);
        } else {
            sl[0] = __fadd_rn(sl[0], p);
        }
        if (k == full - 1) {
#pragma unroll
            for (int q = 0; q < 4; q++) sl[q] = __fadd_rn(sl[q], pl[q]);
        }
    }
    const float a = __fadd_rn(0.f, __fadd_rn(__fadd_rn(sl[0], sl[1]), __fadd_rn(sl[2], sl[3])));
    const bool keep = a >= 0.f;
    if (mask) mask[t] = keep ? 1 : 0;
    y[t] = __fmul_rn(a, keep ? 1.f : 0.f);       // Rectify = a * (a >= 0): -0.0 for negative inputs (nn.go:162-189)
}

// d filter[o][ch][ky][kx] = sum over (b, oy, ox) of dy*mask * x: one thread per filter element and batch slice
__global__ void __launch_bounds__(128)
conv_bwd_filter_kernel(ConvShape c, int n, int nslice, const float *__restrict__ x, const float *__restrict__ dy,
                       const uint8_t *__restrict__ mask, float *__restrict__ gslice, int64_t slice_stride, int64_t f_off)
{
    const int K = c.C * c.KH * c.KW, KK = c.KH * c.KW, S = c.OH * c.OW;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int z = blockIdx.y;
    if (e >= c.O * K) return;
    const int o = e / K, k = e - o * K, ch = k / KK, r = k - ch * KK, ky = r / c.KW, kx = r - ky * c.KW;
    const int per = (n + nslice - 1) / nslice, b0 = z * per, b1 = min(n, b0 + per);
    float acc = 0.f;
    for (int b = b0; b < b1; b++) {
        const float *im = x + ((int64_t)b * c.C + ch) * c.H * c.W;
        const float *d = dy + ((int64_t)b * c.O + o) * S;
        const uint8_t *mk = mask + ((int64_t)b * c.O + o) * S;
        for (int oy = 0; oy < c.OH; oy++) {
            const int iy = -c.pad + ky + oy * c.stride;
            if (iy < 0 || iy >= c.H) continue;
            for (int ox = 0; ox < c.OW; ox++) {
                const int ix = -c.pad + kx + ox * c.stride;
                if (ix < 0 || ix >= c.W) continue;
                const int s = oy * c.OW + ox;
                if (mk[s]) acc += d[s] * __ldg(im + (int64_t)iy * c.W + ix);
            }
        }
    }
    gslice[(int64_t)z * slice_stride + f_off + e] = acc;
}

// dx[b][ch][iy][ix] = sum over (o, ky, kx) with (iy + pad - ky) = oy * stride of dy*mask * f: one thread per input element
__global__ void __launch_bounds__(256)
conv_bwd_data_kernel(ConvShape c, int n, const float *__restrict__ dy, const uint8_t *__restrict__ mask,
                     const float *__restrict__ f, float *__restrict__ dx)
{
    const int S = c.OH * c.OW, KK = c.KH * c.KW, K = c.C * KK;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * c.C * c.H * c.W) return;
    const int ix = (int)(t % c.W), iy = (int)((t / c.W) % c.H), ch = (int)((t / ((int64_t)c.W * c.H)) % c.C);
    const int b = (int)(t / ((int64_t)c.W * c.H * c.C));
    float acc = 0.f;
    for (int ky = 0; ky < c.KH; ky++) {
        const int ny = iy + c.pad - ky;
        if (ny < 0 || ny % c.stride) continue;
        const int oy = ny / c.stride;
        if (oy >= c.OH) continue;
        for (int kx = 0; kx < c.KW; kx++) {
            const int nx = ix + c.pad - kx;
            if (nx < 0 || nx % c.stride) continue;
            const int ox = nx / c.stride;
            if (ox >= c.OW) continue;
            const int s = oy * c.OW + ox;
            for (int o = 0; o < c.O; o++) {
                const int64_t at = ((int64_t)b * c.O + o) * S + s;
                if (mask[at]) acc += dy[at] * __ldg(f + (int64_t)o * K + ch * KK + ky * c.KW + kx);
            }
        }
    }
    dx[t] = acc;
}

}  // namespace

ConvShape make_conv_shape(int C, int H, int W, int O, int KH, int KW, int pad, int stride)
{
    ConvShape c;
    c.C = C; c.H = H; c.W = W; c.O = O; c.KH = KH; c.KW = KW; c.pad = pad; c.stride = stride;
    c.OH = (H + 2 * pad - KH) / stride + 1;      // gorgonia op_nn.go:301-305 (dilation 1)
    c.OW = (W + 2 * pad - KW) / stride + 1;
    return c;
}

AtariNet make_atari_net(int C, int H, int W, int A)
{
    AtariNet t;
    t.c1 = make_conv_shape(C, H, W, 32, 8, 8, 1, 4);                   // policy.go:64
    t.c2 = make_conv_shape(32, t.c1.OH, t.c1.OW, 64, 4, 4, 1, 2);       // policy.go:65
    t.c3 = make_conv_shape(64, t.c2.OH, t.c2.OW, 64, 3, 3, 1, 1);       // policy.go:66
    t.F = 64 * t.c3.OH * t.c3.OW;                                       // Flatten (6400 for 84 x 84)
    t.HID = 512;                                                        // policy.go:68
    t.A = A;
    t.f1 = 0;
    t.f2 = t.f1 + (int64_t)32 * C * 64;
    t.f3 = t.f2 + (int64_t)64 * 32 * 16;
    t.w4 = t.f3 + (int64_t)64 * 64 * 9;
    t.b4 = t.w4 + (int64_t)t.F * t.HID;
    t.w5 = t.b4 + t.HID;
    t.b5 = t.w5 + (int64_t)t.HID * A;
    t.P = t.b5 + A;
    return t;
}

void launch_conv_forward(const ConvShape &c, int n, const float *x, const float *f, float *y, uint8_t *mask, cudaStream_t st)
{
    const int64_t total = (int64_t)n * c.O * c.OH * c.OW;
    conv_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(c, n, x, f, y, mask);
}

void launch_conv_backward_filter(const ConvShape &c, int n, int nslice, const float *x, const float *dy, const uint8_t *mask,
                                 float *gslice, int64_t slice_stride, int64_t f_off, cudaStream_t st)
{
    const int elems = c.O * c.C * c.KH * c.KW;
    dim3 grid((elems + 127) / 128, nslice);
    conv_bwd_filter_kernel<<<grid, 128, 0, st>>>(c, n, nslice, x, dy, mask, gslice, slice_stride, f_off);
}

void launch_conv_backward_data(const ConvShape &c, int n, const float *dy, const uint8_t *mask, const float *f, float *dx,
                               cudaStream_t st)
{
    const int64_t total = (int64_t)n * c.C * c.H * c.W;
    conv_bwd_data_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(c, n, dy, mask, f, dx);
}

}  // namespace gq
