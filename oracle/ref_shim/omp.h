// TEST INFRASTRUCTURE ONLY.  mpc/gait_optimizer.cpp includes <omp.h> for its 10-thread line search; the OpenMP runtime is not
// linkable in this image (no libgomp.spec), so the pragmas are ignored (serial loop) and these calls are no-ops.
#pragma once
inline void omp_set_num_threads(int) {}
inline int omp_get_thread_num() { return 0; }
inline int omp_get_num_threads() { return 1; }
inline int omp_get_max_threads() { return 1; }
