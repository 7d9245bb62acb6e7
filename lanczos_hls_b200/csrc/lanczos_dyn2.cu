// lanczos_dyn2.cu -- second half of the lanczos_dyn_kernel instances (see the dispatch at the end of lanczos_dyn.cu):
// the same source compiled as a second translation unit so that the instances build in parallel.
#define LZD_PART 1
#include "lanczos_dyn.cu"
