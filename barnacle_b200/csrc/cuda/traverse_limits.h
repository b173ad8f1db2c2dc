// Limits shared by the traversal kernel and the host-side scene validation.
#pragma once
namespace bn {
constexpr int kStackSize = 96;     // unified TLAS+BLAS traversal stack entries per ray
constexpr int kMaxLeafCount = 63;  // 6-bit item count in a leaf reference
}  // namespace bn
