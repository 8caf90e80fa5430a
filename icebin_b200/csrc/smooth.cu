// smooth.cu -- sigma smoothing operator  M <- smoothI * M  (slib/icebin/smoother.cpp:8-99,
// RegridMatrices_Dynamic.cpp:236-248).  Filled in below.
#include "types.cuh"

namespace ib {

void smooth_rows(RegridMatrices &rm, Weighted &W, Params const &p)
{
    fail(IB_ERR_ARG, "sigma smoothing is not implemented yet");
}

}   // namespace ib
