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
    std::map<int, double> velx;             // velocity_handleVerts (only the x component is ever non-zero)
    bool has_turn = false;
    int turn_v = 0;
    double turn_lo = 0, turn_hi = 0;

    void init(int kind, int nV, const double* V_rest, double handle_ratio);
    void fixed_mask(uint8_t* out) const;
    void step(double* x, double dt);
};

}  // namespace dotgpu
