// Kernels of the Hyrax opening path (bound, bullet folding, sumcheck rounds).
#pragma once
#include "msm_kernels.cuh"

namespace sbn {
}  // namespace sbn
