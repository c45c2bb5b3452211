// rt_f64.cu -- the FP64 parity build of every kernel.  This translation unit is compiled with -fmad=false so no
// multiply-add is contracted: with the reference's operation order (rt_trace.cuh, M<double>::exact paths) the
// results track the reference's IEEE-double Python arithmetic (north_star: <= 1e-9 relative parity).
#include "rt_kernels.cuh"
#include "rt_wavefront.cuh"
namespace rt {
RT_INSTANTIATE_LAUNCHERS(double)
RT_INSTANTIATE_WAVEFRONT(double)
}
