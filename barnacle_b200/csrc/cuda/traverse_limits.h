// Limits shared by the traversal kernel and the host-side scene validation.
#pragma once
namespace bn {
constexpr int kStackSize = 96;     // unified TLAS+BLAS traversal stack entries per ray
constexpr int kMaxLeafCount = 7;   // 3-bit triangle count in a BLAS leaf reference (the reference builds leaves of <= 4)
}  // namespace bn
