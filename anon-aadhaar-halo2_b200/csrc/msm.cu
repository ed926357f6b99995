// placeholder until the Pippenger kernels land
#include "../../include/b200zk.h"
#include "common.cuh"
namespace zk { void msm_release_bases(Context&) {} }
