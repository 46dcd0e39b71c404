#include "anim_host.h"

#include <cmath>
#include <stdexcept>

#include "../../include/dotgpu.h"

namespace dotgpu {

void AnimHost::init(int kind_, int nV_, const double* V, double ratio) {
    kind = kind_;
    nV = nV_;
    if (kind < DOTGPU_ANIM_NULL || kind > DOTGPU_ANIM_RUBBERBANDPULL) throw std::invalid_argument("unknown anim script");
    double lo[3], hi[3];
    for (int c = 0; c < 3; ++c) lo[c] = hi[c] = V[c];
    for (int v = 1; v < nV; ++v)
        for (int c = 0; c < 3; ++c) {
            lo[c] = std::fmin(lo[c], V[3 * (size_t)v + c]);
            hi[c] = std::fmax(hi[c], V[3 * (size_t)v + c]);
        }
    for (int c = 0; c < 3; ++c) center[c] = (lo[c] + hi[c]) / 2.0;  // bbox.colwise().mean()
    handles.assign(2, {});
    fixed_now.assign(nV, 0);
    if (kind == DOTGPU_ANIM_RUBBERBANDPULL) {
        // AnimScripter.cpp:219-257: bottom and top 2 % of the y range are pulled apart at -/+ 0.2 (handleVerts[1]), the waist
        // (0.48..0.52 of the y range) is dragged in -x at 2.5 (handleVerts[0]) until its first vertex has moved by 5, then released
        const double ry = hi[1] - lo[1];
        bool turning_set = false;
        for (int v = 0; v < nV; ++v) {
            const double y = V[3 * (size_t)v + 1];
            if (y < lo[1] + ry * 0.02) {
                handles[1].push_back(v); velx[v] = 0.0; vely[v] = -0.2;
            } else if (y > hi[1] - ry * 0.02) {
                handles[1].push_back(v); velx[v] = 0.0; vely[v] = 0.2;
            } else if (y < hi[1] - ry * 0.48 && y > lo[1] + ry * 0.48) {
                handles[0].push_back(v); velx[v] = -2.5; vely[v] = 0.0;
                if (!turning_set) {
                    turning_set = true;
                    has_turn = true;
                    turn_v = v;
                    turn_lo = V[3 * (size_t)v] - 5.0;
                }
            }
        }
        if (!turning_set) throw std::invalid_argument("rubberBandPull: no vertex in the waist band");
        for (auto& h : handles)
            for (int v : h) fixed_now[v] = 1;
        return;
    }
    const double range = hi[0] - lo[0];
    for (int v = 0; v < nV; ++v) {
        double xv = V[3 * (size_t)v];
        if (xv < lo[0] + range * ratio) handles[0].push_back(v);
        else if (xv > hi[0] - range * ratio) handles[1].push_back(v);
    }
    double a = 0, vx = 0;
    bool ha = false, hv = false;
    switch (kind) {
        case DOTGPU_ANIM_STRETCH: vx = -0.1; hv = true; break;
        case DOTGPU_ANIM_SQUASH: vx = 0.03; hv = true; break;
        case DOTGPU_ANIM_STRETCHNSQUASH: vx = -0.9; hv = true; break;
        case DOTGPU_ANIM_TWIST: a = -0.1 * M_PI; ha = true; break;
        case DOTGPU_ANIM_TWISTNSTRETCH: a = -0.1 * M_PI; ha = true; vx = -0.1; hv = true; break;
        case DOTGPU_ANIM_TWISTNSNS: a = -0.4 * M_PI; ha = true; vx = -1.2; hv = true; break;
        case DOTGPU_ANIM_TWISTNSNS_OLD: a = -0.4 * M_PI; ha = true; vx = -0.9; hv = true; break;
        default: break;
    }
    for (int b = 0; b < 2; ++b) {
        const double sgn = std::pow(-1.0, b);
        for (int v : handles[b]) {
            if (ha) ang[v] = sgn * a;
            if (hv) velx[v] = sgn * vx;
        }
    }
    if (kind == DOTGPU_ANIM_TWISTNSNS || kind == DOTGPU_ANIM_TWISTNSNS_OLD || kind == DOTGPU_ANIM_STRETCHNSQUASH) {
        if (handles[0].empty()) throw std::invalid_argument("no handle vertices");
        has_turn = true;
        turn_v = handles[0].front();
        turn_lo = V[3 * (size_t)turn_v] - (kind == DOTGPU_ANIM_TWISTNSNS ? 1.2 : 0.8);
        turn_hi = V[3 * (size_t)turn_v] + 0.4;
    }
    if (kind == DOTGPU_ANIM_NULL) {
        fixed_now[0] = 1;  // Mesh::computeFeatures default (Mesh.cpp:593-599)
    } else {
        for (auto& h : handles)
            for (int v : h) fixed_now[v] = 1;
    }
}

void AnimHost::fixed_mask(uint8_t* out) const {
    for (int v = 0; v < nV; ++v) out[v] = fixed_now[v];
}

int AnimHost::step(double* x, double dt) {
    std::vector<double> d(3 * (size_t)nV, 0.0);
    if (kind == DOTGPU_ANIM_RUBBERBANDPULL) {  // AnimScripter.cpp:404-423
        int flag = 0;
        if (!released && x[3 * (size_t)turn_v] <= turn_lo) {
            released = true;
            for (int v : handles[0]) { fixed_now[v] = 0; velx[v] = 0.0; vely[v] = 0.0; }
            for (int v : handles[1]) { velx[v] = 0.0; vely[v] = 0.0; }
            flag = 1;
        }
        for (auto& kv : velx) x[3 * (size_t)kv.first] += 1.0 * (kv.second * dt);
        for (auto& kv : vely) x[3 * (size_t)kv.first + 1] += 1.0 * (kv.second * dt);
        return flag;
    }
    for (auto& kv : ang) {
        // Eigen::AngleAxis(angle, UnitX).toRotationMatrix(), term by term
        const double angle = kv.second * dt, c = std::cos(angle), s = std::sin(angle), c1 = 1.0 - c;
        const double R00 = c1 * 1.0 * 1.0 + c, R11 = c1 * 0.0 * 0.0 + c, R22 = c1 * 0.0 * 0.0 + c, R12 = 0.0 - s, R21 = 0.0 + s;
        const double* xv = x + 3 * (size_t)kv.first;
        const double r0 = xv[0] - center[0], r1 = xv[1] - center[1], r2 = xv[2] - center[2];
        double* dv = d.data() + 3 * (size_t)kv.first;
        dv[0] = (R00 * r0 + center[0]) - xv[0];
        dv[1] = (R11 * r1 + R12 * r2 + center[1]) - xv[1];
        dv[2] = (R21 * r1 + R22 * r2 + center[2]) - xv[2];
    }
    bool flip = false;
    if (has_turn) {
        const double xt = x[3 * (size_t)turn_v];
        flip = (xt <= turn_lo) || (xt >= turn_hi);
    }
    for (auto& kv : velx) {
        if (flip) kv.second *= -1.0;
        d[3 * (size_t)kv.first] += kv.second * dt;
    }
    for (size_t i = 0; i < d.size(); ++i) x[i] += 1.0 * d[i];
    return 0;
}

}  // namespace dotgpu
