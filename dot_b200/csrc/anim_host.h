// Scripted Dirichlet motion on the host: AnimScripter<3> (AnimScripter.cpp:29-453) and
// IglUtils::findBorderVerts (IglUtils.cpp:909-927).
#pragma once
#include <cstdint>
#include <map>
#include <vector>

namespace dotgpu {

struct AnimHost {
    int kind = 0, nV = 0;
    std::vector<std::vector<int>> handles;  // borderVerts_primitive: [0] = x-min side, [1] = x-max side
    double center[3] = {0, 0, 0};           // bbox centre of the rest shape
    std::map<int, double> ang;              // angVel_handleVerts (std::map order, as iterated by the reference)
    std::map<int, double> velx;             // velocity_handleVerts, x component (the bar scripts)
    std::map<int, double> vely;             // velocity_handleVerts, y component (rubberBandPull: bottom / top pulled apart)
    std::vector<uint8_t> fixed_now;         // the current Dirichlet set (rubberBandPull releases the waist handle mid-run)
    bool released = false;
    bool has_turn = false;
    int turn_v = 0;
    double turn_lo = 0, turn_hi = 0;

    void init(int kind, int nV, const double* V_rest, double handle_ratio);
    void fixed_mask(uint8_t* out) const;
    // returns 1 when the Dirichlet set changed in this step (AnimScripter::stepAnimScript's return value -> the caller runs
    // updatePrecondMtrAndFactorize = dotgpu_stepper_set_fixed, Optimizer.cpp:334-336), else 0
    int step(double* x, double dt);
};

}  // namespace dotgpu
