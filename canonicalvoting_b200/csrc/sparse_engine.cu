// canonicalvoting_b200/csrc/sparse_engine.cu -- inference executor of the sparse-voxel U-Net + head decode.
//
// The reference's eval path runs utils/minkunet.py:122-180 layer by layer from Python: per convolution a
// MinkowskiEngine call, then BatchNorm, ReLU, residual add and ME.cat as separate torch kernels
// (eval_joint.py:169-171), followed by ~10 small torch ops for the head decode (eval_joint.py:173-190).
// Here the host side compiles the network ONCE into a flat program of fused convolution ops
//     out[:, c0:c0+cout] = [relu]( conv(in[:, a0:a0+cin]) + bias + residual )
// (BatchNorm folded into weights / bias, ME.cat realised by writing into column slices of a shared buffer)
// and this file launches the whole program back to back from C++ -- no interpreter between kernels.
#include "common.cuh"

namespace cvb200 {

int launch_conv_persist(const float *d_in, int64_t n_in, int ldi, int cin, const float *d_wt, int cout, const int32_t *d_nbr,
                        int64_t n_out, int k3, const float *d_bias, const float *d_res, int ldr, int relu, float *d_out, int ldo,
                        cudaStream_t stream, int g4, const int32_t *d_n_out, const cvb200_decode_args *decode = nullptr);

// im2col of a small-width input: col[o, k*cin + c] = in[nbr[o,k], c] (0 where the neighbour is missing; the columns
// beyond K^3*cin up to ldo are zero padding).  One thread per (row, offset): a warp writes 32*cin consecutive floats.
// With it the 3-channel 5^3 stem (375 -> 384 columns) becomes a plain [N,384] x [384,32] product on the tensor cores.
__global__ void sc_im2col_kernel(const float *__restrict__ in, int ldi, int cin, const int *__restrict__ nbr, int n_out, int k3,
                                 float *__restrict__ col, int ldo) {
    const int kslots = ldo / cin;                     // >= k3; slots k3..kslots-1 are padding
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_out * kslots) return;
    const int o = (int)(t / kslots), k = (int)(t - (long long)o * kslots);
    const int idx = k < k3 ? __ldg(nbr + (size_t)o * k3 + k) : -1;
    float *dst = col + (size_t)o * ldo + k * cin;
    for (int c = 0; c < cin; c++) dst[c] = idx >= 0 ? __ldg(in + (size_t)idx * ldi + c) : 0.f;
}

// Head decode of the joint model (eval_joint.py:173-190): one thread per point.
//   feats [n, 6*C + C + 1]: xyz[C][3] | scale[C][3] | class logits [C+1] (last = background)
// With `coords` (int32 rows (batch, x, y, z)) it also writes the vote op's first argument, points = coords[:, 1:] * res
// (eval_joint.py:193), so that no torch kernel sits between the network and hv_cuda.forward.
__global__ void head_decode_kernel(const float *__restrict__ f, int ld, int n, int nclasses, int log_scale,
                                   float *__restrict__ xyz, float *__restrict__ scale, long long *__restrict__ cls,
                                   float *__restrict__ prob, const int4 *__restrict__ coords, float res, float *__restrict__ points) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (coords) {
        const int4 c = __ldg(coords + i);
        points[3 * (size_t)i] = __fmul_rn((float)c.y, res);
        points[3 * (size_t)i + 1] = __fmul_rn((float)c.z, res);
        points[3 * (size_t)i + 2] = __fmul_rn((float)c.w, res);
    }
    const float *row = f + (size_t)i * ld;
    const float *logit = row + 6 * nclasses;
    float best = -INFINITY, best_obj = -INFINITY;
    int k = 0, k_obj = 0;
    for (int c = 0; c <= nclasses; c++) {          // first maximum, like torch.argmax
        const float v = __ldg(logit + c);
        if (v > best) { best = v; k = c; }
        if (c < nclasses && v > best_obj) { best_obj = v; k_obj = c; }
    }
    float sum = 0.f;
    for (int c = 0; c <= nclasses; c++) sum += expf(__ldg(logit + c) - best);
    if (k == nclasses) k = 0;                      // class_label_idx[class_label_idx == nclasses] = 0  (:178)
#pragma unroll
    for (int d = 0; d < 3; d++) {
        xyz[3 * (size_t)i + d] = __ldg(row + 3 * k + d);
        const float s = __ldg(row + 3 * nclasses + 3 * k + d);
        scale[3 * (size_t)i + d] = log_scale ? expf(s) : s;
    }
    cls[i] = k_obj;                                // argmax over the object classes          (:188)
    prob[i] = expf(best_obj - best) / sum;         // max softmax over the object classes     (:189)
}

}  // namespace cvb200

using namespace cvb200;

extern "C" int cvb200_sc_run_program(const cvb200_sc_op *ops, int32_t n_ops, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVB_REQUIRE(ops && n_ops >= 0, CVB200_EINVAL, "sc_run_program: NULL program");
    for (int i = 0; i < n_ops; i++) {
        const cvb200_sc_op &o = ops[i];
        if (o.kind == CVB200_OP_CONV_TC) {
            const int rc = launch_conv_persist(o.in, o.n_in, o.ldi, o.cin, o.w, o.cout, o.table, o.n_out, o.k3, o.bias, o.residual, o.ldr,
                                               o.relu, o.out, o.ldo, stream, 0, o.n_out_dev, o.decode);
            if (rc) return rc;
        } else if (o.kind == CVB200_OP_CONV_TC_GATHER4) {
            // 4-channel input gathered 8 neighbours per k-block: o.k3 = table width, o.cin = 32 * ceil(k3 / 8) = K of the weights
            const int rc = launch_conv_persist(o.in, o.n_in, o.ldi, o.cin, o.w, o.cout, o.table, o.n_out, 1, o.bias, o.residual, o.ldr, o.relu,
                                               o.out, o.ldo, stream, o.k3, o.n_out_dev);
            if (rc) return rc;
        } else if (o.kind == CVB200_OP_IM2COL) {
            CVB_REQUIRE(!o.n_out_dev, CVB200_EINVAL, "sc_run_program: op %d: a device-side row count needs a tensor-core kind", i);
            CVB_REQUIRE(o.cin >= 1 && o.cin <= 8 && o.ldo % o.cin == 0 && o.ldo >= o.k3 * o.cin && o.in && o.out && o.table, CVB200_EINVAL,
                        "sc_run_program: op %d: im2col needs cin <= 8 and ldo a multiple of cin covering K^3*cin", i);
            if (o.n_out > 0) {
                const long long total = (long long)o.n_out * (o.ldo / o.cin);
                sc_im2col_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(o.in, o.ldi, o.cin, o.table, (int)o.n_out, o.k3, o.out,
                                                                                   o.ldo);
                CVB_LAUNCH_CHECK("sc_im2col_kernel");
            }
        } else {
            CVB_REQUIRE(false, CVB200_EINVAL, "sc_run_program: op %d has unknown kind %d", i, o.kind);
        }
    }
    return 0;
}

extern "C" int cvb200_head_decode_points(const float *d_feats, int32_t ld, int64_t n, int32_t nclasses, int32_t log_scale, float *d_xyz,
                                         float *d_scale, int64_t *d_class, float *d_prob, const int32_t *d_coords, float res,
                                         float *d_points, void *stream_) {
    CVB_REQUIRE(n >= 0 && n < (1LL << 31) && nclasses >= 1 && nclasses <= 64 && ld >= 7 * nclasses + 1, CVB200_EINVAL,
                "head_decode: bad sizes (n=%lld, nclasses=%d, ld=%d)", (long long)n, nclasses, ld);
    if (n == 0) return 0;
    CVB_REQUIRE(d_feats && d_xyz && d_scale && d_class && d_prob && (d_coords == nullptr) == (d_points == nullptr), CVB200_EINVAL,
                "head_decode: NULL argument (coords and points go together)");
    head_decode_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream_>>>(d_feats, ld, (int)n, nclasses, log_scale, d_xyz,
                                                                                     d_scale, (long long *)d_class, d_prob,
                                                                                     (const int4 *)d_coords, res, d_points);
    CVB_LAUNCH_CHECK("head_decode_kernel");
    return 0;
}

extern "C" int cvb200_head_decode(const float *d_feats, int32_t ld, int64_t n, int32_t nclasses, int32_t log_scale, float *d_xyz,
                                  float *d_scale, int64_t *d_class, float *d_prob, void *stream_) {
    CVB_REQUIRE(n >= 0 && n < (1LL << 31) && nclasses >= 1 && nclasses <= 64 && ld >= 7 * nclasses + 1, CVB200_EINVAL,
                "head_decode: bad sizes (n=%lld, nclasses=%d, ld=%d)", (long long)n, nclasses, ld);
    if (n == 0) return 0;
    CVB_REQUIRE(d_feats && d_xyz && d_scale && d_class && d_prob, CVB200_EINVAL, "head_decode: NULL argument");
    head_decode_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream_>>>(d_feats, ld, (int)n, nclasses, log_scale, d_xyz,
                                                                                     d_scale, (long long *)d_class, d_prob, nullptr, 0.f, nullptr);
    CVB_LAUNCH_CHECK("head_decode_kernel");
    return 0;
}
