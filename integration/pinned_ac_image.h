// pinned_ac_image.h -- the coefficient storage class a libjxl host installs for a GPU frame
// (INTEGRATION.md §2).  It implements libjxl's abstract jxl::ACImage (lib/jxl/dct_util.h:41-54), so the
// UNMODIFIED entropy decoder writes each group's coefficients through PlaneRow(c, group, offset)
// (lib/jxl/dec_group.cc:335-338) straight into one page-locked allocation laid out
//
//     [group][channel X,Y,B][65536]
//
// which is the layout jxlgpu_submit_group(s) turns into one DMA per group (or per run of groups).
// The allocator is a pair of function pointers: jxlgpu_alloc_pinned / jxlgpu_free_pinned in
// production, malloc / free in the CPU test that drives the reference decoder through this class
// (tests/test_oracle_vs_reference.py::test_group_major_ac_image_is_a_drop_in).
// This header is ours; it only includes the reference's interface header.
#ifndef JXL_B200_INTEGRATION_PINNED_AC_IMAGE_H_
#define JXL_B200_INTEGRATION_PINNED_AC_IMAGE_H_

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>

#include "lib/jxl/dct_util.h"

namespace jxlb_integration {

template <typename T>
class GroupMajorACImage final : public jxl::ACImage {
 public:
  using AllocFn = void* (*)(size_t);
  using FreeFn = void (*)(void*);
  static constexpr size_t kGroupCoeffs = 65536;  // kGroupDim * kGroupDim, frame_dimensions.h:25

  static std::unique_ptr<GroupMajorACImage> Make(size_t num_groups, AllocFn alloc, FreeFn release) {
    std::unique_ptr<GroupMajorACImage> im(new GroupMajorACImage());
    im->bytes_ = num_groups * 3 * kGroupCoeffs * sizeof(T);
    im->base_ = num_groups ? static_cast<T*>(alloc(im->bytes_)) : nullptr;
    im->release_ = release;
    im->num_groups_ = num_groups;
    if (num_groups && !im->base_) return nullptr;
    return im;
  }
  ~GroupMajorACImage() override {
    if (base_) release_(base_);
  }

  jxl::ACType Type() const override { return sizeof(T) == 2 ? jxl::ACType::k16 : jxl::ACType::k32; }
  // c = channel, y = AC group index, xbase = coefficient offset inside the group
  jxl::ACPtr PlaneRow(size_t c, size_t y, size_t xbase) override {
    return jxl::ACPtr(base_ + (y * 3 + c) * kGroupCoeffs + xbase);
  }
  jxl::ConstACPtr PlaneRow(size_t c, size_t y, size_t xbase) const override {
    return jxl::ConstACPtr(static_cast<const T*>(base_) + (y * 3 + c) * kGroupCoeffs + xbase);
  }
  size_t PixelsPerRow() const override { return 3 * kGroupCoeffs; }
  void ZeroFill() override {
    if (base_) memset(base_, 0, bytes_);
  }
  void ZeroFillPlane(size_t c) override {
    for (size_t g = 0; g < num_groups_; g++) memset(base_ + (g * 3 + c) * kGroupCoeffs, 0, kGroupCoeffs * sizeof(T));
  }
  bool IsEmpty() const override { return num_groups_ == 0; }

  // what jxlgpu_submit_group(s) is given for group g: coeff[c] = GroupBlock(g) + c * 65536
  const T* GroupBlock(size_t g) const { return base_ + g * 3 * kGroupCoeffs; }
  void* data() { return base_; }
  size_t size_bytes() const { return bytes_; }

 private:
  GroupMajorACImage() = default;
  T* base_ = nullptr;
  size_t bytes_ = 0, num_groups_ = 0;
  FreeFn release_ = nullptr;
};

}  // namespace jxlb_integration
#endif  // JXL_B200_INTEGRATION_PINNED_AC_IMAGE_H_
