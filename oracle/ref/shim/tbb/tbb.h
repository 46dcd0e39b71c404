// TEST INFRASTRUCTURE (oracle build only).
// Minimal stand-in for the one TBB entry point the reference uses
// (tbb::parallel_for(first,last,step,body), 96 call sites).
// Every loop body in the reference writes disjoint outputs, so an OpenMP static
// schedule gives bit-identical results for any thread count.
#pragma once
namespace tbb {
template <class I, class F>
inline void parallel_for(I first, I last, I step, const F& f) {
#pragma omp parallel for schedule(static)
    for (I i = first; i < last; i += step) f(i);
}
}  // namespace tbb
