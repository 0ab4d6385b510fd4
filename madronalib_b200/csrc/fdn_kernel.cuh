// fdn_kernel.cuh -- fused 3-operator FM -> FDN<8> kernel (config 4).  Placeholder until the
// specialised kernel lands: graphs of this shape run through the generic interpreter.
#pragma once
#include <vector>

#include "ops.cuh"

namespace mlb
{
struct FdnArgs
{
  int dummy;
};

inline bool match_fm3_fdn8(const std::vector<mlb_node>&, const std::vector<int32_t>&,
                           const std::vector<int32_t>&, const std::vector<int32_t>&, FdnArgs*)
{
  return false;
}

inline int launch_fm3_fdn8(const FdnArgs&, bool, uint32_t*, const float*, float*, float*, int,
                           long long, const float*, float*, float*, int, int, int, int,
                           cudaStream_t)
{
  return MLB_ERR_UNSUPPORTED;
}
}  // namespace mlb
