/* jxl_oracle.h -- CPU restatement of the reference hot path. TEST INFRASTRUCTURE ONLY
 * (see jxl_oracle.c). Takes the very same jxlgpu_frame description the product ABI takes. */
#ifndef JXL_ORACLE_H_
#define JXL_ORACLE_H_
#include <stddef.h>
#include <stdint.h>

#include "../include/jxl_b200.h"

#ifdef __cplusplus
extern "C" {
#endif
int jxo_covered_blocks_x(int strategy);
int jxo_covered_blocks_y(int strategy);
/* jxl::TransformToPixels (dec_transforms-inl.h:456-689); coeffs are not modified. */
int jxo_transform_to_pixels(int strategy, const float* coeffs, float* pixels, size_t stride);
/* jxl::LowestFrequenciesFromDC (dec_transforms-inl.h:691-818); writes the LLF corner of block. */
int jxo_llf_from_dc(int strategy, const float* dc, size_t dc_stride, float* block);
/* ComputeScaledDCT<rows,cols> (dct-inl.h:349-371). */
int jxo_scaled_dct(int rows, int cols, const float* px, size_t stride, float* out);
/* AdjustQuantBias (quantizer-inl.h:35-67); rcp_mode 0 exact reciprocal, 1 host rcpss. */
float jxo_adjust_quant_bias(int c, int32_t q, const float* biases, int rcp_mode);
uint32_t jxo_effective_stage_mask(const jxlgpu_frame* f);
/* ComputeSigma (epf.cc:39-133): (ysize_blocks+4) x (xsize_blocks+4) inverse sigmas. */
void jxo_compute_sigma(const jxlgpu_frame* f, float* sigma);
/* DC stage in front of the path: DequantDC 4:4:4 (compressed_dc.cc:199-232) and
 * AdaptiveDCSmoothing (:50-197); planes [3][ys][xs], X/Y/B order. */
void jxo_dequant_dc(const int32_t* const q[3], size_t xs, size_t ys, const float* dc_factors, float mul,
                    const float* cfl_factors, float* out);
int jxo_adaptive_dc_smoothing(const float* dc_factors, float* dc, size_t xs, size_t ys);
/* TF_SRGB::EncodedFromDisplay (cms/transfer_functions-inl.h:244-267). */
float jxo_srgb_from_linear(float v);
/* MakeUnsigned (stage_write.cc:455-479), bits = 8 (dithered) or 16. */
uint32_t jxo_make_unsigned(float v, int bits, size_t x, size_t y, int c);
/* float -> binary16, round to nearest even (stage_write.cc:590-640). */
uint16_t jxo_f16_from_f32(float v);
size_t jxo_out_bytes(const jxlgpu_frame* f);
/* Whole frame: coeff[c] = [num_groups][65536] host planes of f->ac_type.
 * out: jxo_out_bytes(f) bytes, dense rows, in f->out_format. Returns 0 on success. */
int jxo_render_frame(const jxlgpu_frame* f, const void* const coeff[3], int rcp_mode, void* out);
/* UpsamplingStage (stage_upsampling.cc:51-271): weights -> N*N*25 kernel; one plane w x h -> ow x oh */
void jxo_upsampling_kernel(int N, const float* weights, float* kernel);
void jxo_upsample_plane(int N, const float* kernel, const float* in, size_t w, size_t h, size_t ps_in, float* out,
                        size_t ow, size_t oh, size_t ps_out);
#ifdef __cplusplus
}
#endif
#endif
