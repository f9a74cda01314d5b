// gpu_frame_binding.h -- fills the jxlgpu_frame of include/jxl_b200.h from libjxl's decoder state: the
// table of INTEGRATION.md §1 as code.  Called by FrameDecoder after ProcessACGlobal
// (lib/jxl/dec_frame.cc:693-696).  Everything is passed by pointer + stride into libjxl's own images
// (no copies) except the AC strategy plane, whose raw bytes have no public accessor in the unmodified
// tree (a `RawRow()` next to AcStrategyImage::ConstRow, lib/jxl/ac_strategy.h:244-262, would remove
// that copy too).  This header is ours; it only includes reference headers.
// Exercised on the CPU against the unmodified reference: oracle/ref_harness.cc:ref_frame_bind_gpu_frame
// + tests/test_oracle_vs_reference.py::test_gpu_frame_binding_from_decoder_state.
#ifndef JXL_B200_INTEGRATION_GPU_FRAME_BINDING_H_
#define JXL_B200_INTEGRATION_GPU_FRAME_BINDING_H_

#include <cstdint>
#include <cstring>
#include <vector>

#include "jxl_b200.h"
#include "lib/jxl/ac_strategy.h"
#include "lib/jxl/dec_cache.h"
#include "lib/jxl/frame_header.h"
#include "lib/jxl/loop_filter.h"
#include "lib/jxl/passes_state.h"
#include "lib/jxl/quant_weights.h"

namespace jxlb_integration {

// Frames the GPU path takes; everything else stays on libjxl's CPU path.
inline bool IsEligible(const jxl::FrameHeader& fh, const jxl::CodecMetadata& metadata) {
  using jxl::FrameHeader;
  // XYB frames, or JPEG-origin frames (YCbCr colour transform on a non-XYB image) without chroma subsampling
  const bool xyb = metadata.m.xyb_encoded && fh.color_transform == jxl::ColorTransform::kXYB;
  const bool ycbcr = !metadata.m.xyb_encoded && fh.color_transform == jxl::ColorTransform::kYCbCr;
  return fh.encoding == jxl::FrameEncoding::kVarDCT && (xyb || ycbcr) && fh.chroma_subsampling.Is444() &&
         (fh.upsampling == 1 || fh.upsampling == 2 || fh.upsampling == 4 || fh.upsampling == 8) &&
         !(fh.flags & (FrameHeader::kPatches | FrameHeader::kSplines)) &&   // (kNoise: generated and added on the device)
         metadata.m.num_extra_channels == 0;   // (several passes: accumulated in the dense pinned storage)
}

struct GpuFrameBinding {
  jxlgpu_frame frame;
  std::vector<uint8_t> ac_strategy;  // (type << 1) | is_first, see the header comment
};

// out_format / stage_mask: what the application asked for (JxlPixelFormat, output colour encoding);
// see INTEGRATION.md §1 for the mapping.
inline bool BindGpuFrame(const jxl::PassesDecoderState& ds, const jxl::FrameHeader& fh, uint32_t out_format,
                         uint32_t stage_mask, GpuFrameBinding* b) {
  const jxl::PassesSharedState& sh = *ds.shared;
  const jxl::FrameDimensions& d = sh.frame_dim;
  jxlgpu_frame& f = b->frame;
  memset(&f, 0, sizeof(f));
  f.xsize = static_cast<uint32_t>(d.xsize);
  f.ysize = static_cast<uint32_t>(d.ysize);
  f.xsize_blocks = static_cast<uint32_t>(d.xsize_blocks);
  f.ysize_blocks = static_cast<uint32_t>(d.ysize_blocks);
  f.ac_type = ds.coefficients->Type() == jxl::ACType::k16 ? JXLGPU_AC_INT16 : JXLGPU_AC_INT32;

  const size_t xb = d.xsize_blocks, yb = d.ysize_blocks;
  b->ac_strategy.resize(xb * yb);
  for (size_t y = 0; y < yb; y++) {
    jxl::AcStrategyRow row = sh.ac_strategy.ConstRow(y);
    for (size_t x = 0; x < xb; x++)
      b->ac_strategy[y * xb + x] = static_cast<uint8_t>((row[x].RawStrategy() << 1) | (row[x].IsFirstBlock() ? 1 : 0));
  }
  f.ac_strategy = b->ac_strategy.data();
  f.ac_strategy_stride = xb;
  f.raw_quant = sh.raw_quant_field.ConstRow(0);
  f.raw_quant_stride = sh.raw_quant_field.PixelsPerRow();
  if (fh.loop_filter.epf_iters > 0) {
    f.epf_sharpness = sh.epf_sharpness.ConstRow(0);
    f.epf_sharpness_stride = sh.epf_sharpness.PixelsPerRow();
  }
  f.ytox_map = sh.cmap.ytox_map.ConstRow(0);
  f.ytob_map = sh.cmap.ytob_map.ConstRow(0);
  f.cmap_stride = sh.cmap.ytox_map.PixelsPerRow();
  if (sh.cmap.ytob_map.PixelsPerRow() != f.cmap_stride) return false;
  for (size_t c = 0; c < 3; c++) f.dc[c] = sh.dc->ConstPlaneRow(c, 0);
  f.dc_stride = sh.dc->PixelsPerRow();

  // dequantisation matrices: one table, offsets relative to its lowest address (quant_weights.h:364-367)
  const float* base = sh.matrices.Matrix(jxl::AcStrategyType::DCT, 0);
  for (size_t k = 0; k < jxl::AcStrategy::kNumValidStrategies; k++)
    for (size_t c = 0; c < 3; c++) {
      const float* p = sh.matrices.Matrix(static_cast<jxl::AcStrategyType>(k), c);
      if (p < base) base = p;
    }
  f.dequant_table = base;
  f.dequant_table_floats = jxl::DequantMatrices::kSumRequiredXy * jxl::kDCTBlockSize * 3;
  for (size_t k = 0; k < jxl::AcStrategy::kNumValidStrategies; k++)
    for (size_t c = 0; c < 3; c++)
      f.dequant_offsets[k * 3 + c] =
          static_cast<uint32_t>(sh.matrices.Matrix(static_cast<jxl::AcStrategyType>(k), c) - base);

  f.inv_global_scale = sh.quantizer.InvGlobalScale();
  f.quant_scale = sh.quantizer.Scale();
  f.x_dm_multiplier = ds.x_dm_multiplier;
  f.b_dm_multiplier = ds.b_dm_multiplier;
  const jxl::OpsinParams& op = ds.output_encoding_info.opsin_params;
  memcpy(f.quant_biases, op.quant_biases, sizeof(f.quant_biases));
  const jxl::ColorCorrelation& cc = sh.cmap.base();
  f.cfl_base_x = cc.GetBaseCorrelationX();
  f.cfl_base_b = cc.GetBaseCorrelationB();
  f.cfl_color_scale = 1.0f / static_cast<uint32_t>(cc.GetColorFactor());  // chroma_from_luma.h:77-79

  const jxl::LoopFilter& lf = fh.loop_filter;
  f.gab = lf.gab ? 1 : 0;
  f.gab_weights[0] = lf.gab_x_weight1; f.gab_weights[1] = lf.gab_x_weight2;
  f.gab_weights[2] = lf.gab_y_weight1; f.gab_weights[3] = lf.gab_y_weight2;
  f.gab_weights[4] = lf.gab_b_weight1; f.gab_weights[5] = lf.gab_b_weight2;
  f.epf_iters = lf.epf_iters;
  memcpy(f.epf_sharp_lut, lf.epf_sharp_lut, sizeof(f.epf_sharp_lut));
  memcpy(f.epf_channel_scale, lf.epf_channel_scale, sizeof(f.epf_channel_scale));
  f.epf_quant_mul = lf.epf_quant_mul;
  f.epf_pass0_sigma_scale = lf.epf_pass0_sigma_scale;
  f.epf_pass2_sigma_scale = lf.epf_pass2_sigma_scale;
  f.epf_border_sad_mul = lf.epf_border_sad_mul;

  for (int i = 0; i < 9; i++) f.inverse_opsin_matrix[i] = op.inverse_opsin_matrix[i * 4];  // dec_xyb.h:28-34
  for (int i = 0; i < 3; i++) {
    f.opsin_biases[i] = op.opsin_biases[i];
    f.opsin_biases_cbrt[i] = op.opsin_biases_cbrt[i];
  }
  f.out_format = out_format;
  f.stage_mask = stage_mask;
  f.color_transform = fh.color_transform == jxl::ColorTransform::kYCbCr ? 1u : 0u;   // kYCbCrStage instead of XYBStage
  if (fh.flags & jxl::FrameHeader::kNoise) {  // ConvolveNoise + AddNoise (dec_cache.cc:232-236), seeds: dec_cache.h:127-128
    f.noise = 1;
    for (size_t i = 0; i < 8; i++) f.noise_lut[i] = sh.image_features.noise_params.lut[i];
    f.visible_frame_index = static_cast<uint32_t>(ds.visible_frame_index);
    f.nonvisible_frame_index = static_cast<uint32_t>(ds.nonvisible_frame_index);
  }
  if (fh.upsampling != 1) {  // UpsamplingStage of the colour channels (dec_cache.cc:216-227)
    const jxl::CustomTransformData& td = fh.nonserialized_metadata->transform_data;
    f.upsampling = fh.upsampling;
    f.xsize_upsampled = static_cast<uint32_t>(d.xsize_upsampled);
    f.ysize_upsampled = static_cast<uint32_t>(d.ysize_upsampled);
    f.upsampling_weights = fh.upsampling == 2 ? td.upsampling2_weights
                           : fh.upsampling == 4 ? td.upsampling4_weights : td.upsampling8_weights;
  }
  return true;
}

}  // namespace jxlb_integration
#endif  // JXL_B200_INTEGRATION_GPU_FRAME_BINDING_H_
