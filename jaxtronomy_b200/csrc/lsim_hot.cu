// lsim_hot.cu -- explicit instantiations of the headline step's kernels (see lsim_hot_list.h for why they live in a
// translation unit of their own).  No host code: lsim_api.cu launches them through their `extern template` declarations.
#include "../../include/lsim.h"
#include "lsim_kernels.cuh"

#define LSIM_HOT template
#include "lsim_hot_list.h"
#undef LSIM_HOT
