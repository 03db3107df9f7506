// tc_kernels.h -- host handles of the tcgen05 kernels that are not plain convolutions
#pragma once
#include "kernels.h"

namespace sga {

// igdn_tc.cu
struct IgdnTc;
int igdn_tc_create(const __nv_bfloat16* gamma_op, int C, long long pixels, int bwd, IgdnTc** out);
void igdn_tc_destroy(IgdnTc* c);
int igdn_tc_forward(const IgdnTc* c, const __nv_bfloat16* x, const float* beta, __nv_bfloat16* out, __nv_bfloat16* s_out,
                    cudaStream_t st);
// x_is_act: `x` holds the saved activation act = x*s instead of x (the fused forward kernels do not store x)
int igdn_tc_backward(const IgdnTc* c, const __nv_bfloat16* g, const __nv_bfloat16* x, const __nv_bfloat16* s,
                     __nv_bfloat16* gx, cudaStream_t st, int x_is_act = 0);

// layer3_tc.cu
struct L3Tc;
int l3_pack_weights(const float* k3, int C, __nv_bfloat16* w_scatter, __nv_bfloat16* w_gather, cudaStream_t st);
// gamma_bwd_op / x_saved / s_saved: optional; bf16 [C][C] K-major gamma of IGDN_2 and its saved tensors [N,Hin,Win,C],
// enabling l3_tc_dgrad_igdn
int l3_tc_create(const __nv_bfloat16* act, const __nv_bfloat16* w_scatter, const __nv_bfloat16* w_gather,
                 const __nv_bfloat16* gamma_bwd_op, const __nv_bfloat16* x_saved, const __nv_bfloat16* s_saved, int N,
                 int Hin, int Win, int C, int H, int W, L3Tc** out);
void l3_tc_destroy(L3Tc* c);
int l3_tc_forward(const L3Tc* c, const float* bias_host3, const float* x, float grad_scale, float* xt_out,
                  __nv_bfloat16* gx, double* sse, double* sse_r, cudaStream_t st);
int l3_tc_dgrad(const L3Tc* c, const __nv_bfloat16* gx, __nv_bfloat16* out, cudaStream_t st);
// dgrad fused with the backward of the IGDN in front of layer 3: out = d loss / d x_2 (needs C in {64,128,192})
bool l3_tc_can_fuse_igdn(const L3Tc* c);
int l3_tc_dgrad_igdn(const L3Tc* c, const __nv_bfloat16* gx, __nv_bfloat16* out, cudaStream_t st);

}  // namespace sga
