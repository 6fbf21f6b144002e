/* Stand-in for <claraparabricks/genomeworks/cudaaligner/cudaaligner.hpp>: cudapolisher.cpp:45 calls cudaaligner::Init(). */
#pragma once
namespace claraparabricks {
namespace genomeworks {
namespace cudaaligner {
inline void Init() {}
}  // namespace cudaaligner
}  // namespace genomeworks
}  // namespace claraparabricks
