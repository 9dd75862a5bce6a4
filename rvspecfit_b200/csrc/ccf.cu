// placeholder: CCF kernels; filled in next
#include "common.cuh"
