// lsim_hot_list.h -- the kernels of the headline step, instantiated in their own translation unit (lsim_hot.cu).
//
// Why: liblsim.so's main translation unit (lsim_api.cu) carries ~150 kernel instantiations and is compiled with
// `--split-compile`, which partitions the module before optimisation; the partitioning changes the code ptxas produces
// for individual kernels (measured in round 2: k_likelihood<1> came out at 48 registers and 0.33 ms instead of 64 and
// 0.29 ms, k_image_fold gained spills in its row loop, depending on what else was in the module).  The hot kernels are
// therefore compiled alone, without splitting; lsim_api.cu sees them as `extern template`.
//
// LSIM_HOT(...) expands to `template` (lsim_hot.cu) or `extern template` (lsim_api.cu).
#pragma once

#define LSIM_HOT_FOLD(NS, SS, NL)                                                                                      \
  LSIM_HOT __global__ void k_image_fold<NS, SS, NL>(PlanDev, const double*, int64_t, int, double, double*,              \
                                                     const unsigned char*, int);
#define LSIM_HOT_LIKE(NS)                                                                                              \
  LSIM_HOT __global__ void k_likelihood<NS>(PlanDev, const double*, int64_t, int, int, double*, double*, double*,       \
                                             const double*);
#define LSIM_HOT_CONV(KT)                                                                                              \
  LSIM_HOT __global__ void k_conv_tiled_u<KT, 15, 4>(const double*, double*, int, int, int, double, int,                \
                                                      const __grid_constant__ ConvWeights<KT>);                                          \
  LSIM_HOT __global__ void k_conv_tiled_u<KT, 10, 6>(const double*, double*, int, int, int, double, int,                \
                                                      const __grid_constant__ ConvWeights<KT>);                                          \
  LSIM_HOT __global__ void k_conv_tiled_u<KT, 8, 8>(const double*, double*, int, int, int, double, int,                 \
                                                     const __grid_constant__ ConvWeights<KT>);                                           \
  LSIM_HOT __global__ void k_conv_tiled_fxy<KT, 12, 5>(const double*, double*, int, int, int, double, int,              \
                                                        const __grid_constant__ ConvWeightsFXY<KT>);                   \
  LSIM_HOT __global__ void k_conv_tiled_fxy<KT, 16, 4>(const double*, double*, int, int, int, double, int,              \
                                                        const __grid_constant__ ConvWeightsFXY<KT>);

namespace lsim {
LSIM_HOT_FOLD(1, 3, 1) LSIM_HOT_FOLD(1, 3, 2) LSIM_HOT_FOLD(2, 3, 2)
LSIM_HOT_FOLD(1, 2, 1) LSIM_HOT_FOLD(1, 2, 2) LSIM_HOT_FOLD(2, 2, 2)
LSIM_HOT_FOLD(1, 1, 1) LSIM_HOT_FOLD(1, 1, 2) LSIM_HOT_FOLD(2, 1, 2)
LSIM_HOT_LIKE(1) LSIM_HOT_LIKE(2) LSIM_HOT_LIKE(4) LSIM_HOT_LIKE(8)
LSIM_HOT_CONV(7) LSIM_HOT_CONV(11) LSIM_HOT_CONV(15) LSIM_HOT_CONV(21) LSIM_HOT_CONV(25) LSIM_HOT_CONV(31)
}  // namespace lsim

#undef LSIM_HOT_FOLD
#undef LSIM_HOT_LIKE
#undef LSIM_HOT_CONV
