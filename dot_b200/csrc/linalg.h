// Matrix fill from elemental Hessians, symmetric SpMV / quadratic form on CSR-upper storage, and the
// dense-vector kernels of the L-BFGS iteration (all scalars stay on the device).
#pragma once
#include "comm.h"
#include "common.h"

namespace dotgpu {

// ---- matrix fill (DOTTimeStepper.cpp:574-616, 619-797 as ordered gather lists) ----
struct DeviceFill {
    long long nblk = 0;
    DevBuf<long long> ptr;   // [nblk+1]
    DevBuf<int> src;         // codes, see mesh_host.h
    DevBuf<int> row;         // [nblk] index of scalar row 3v in the concatenated ia array
    DevBuf<int> j;           // [nblk] block index within the row, -1 => fixed vertex (diagonal-only row)
    DevBuf<double> consts;
    DevBuf<int> ia;          // concatenated row pointers, already offset into the concatenated value array
};
void launch_fill(const DeviceFill& f, const double* He, double* a, cudaStream_t st);

// ---- CSR-upper symmetric matrix (global Hessian, Optimizer::linSysSolver) ----
// out[0] = p^T A p  (device scalar), deterministic
void launch_quadform(int n, const int* ia, const int* ja, const double* a, const double* p, double* partial, double* out,
                     cudaStream_t st);
// y = A x with A symmetric given by its upper triangle; needs the transposed index (tp, ti = slot, tr = row)
void launch_spmv_sym(int n, const int* ia, const int* ja, const double* a, const int* tp, const int* tslot, const int* trow,
                     const double* x, double* y, cudaStream_t st);

// ---- dense vectors; scalars live in a small device array `sc` ----
int dot_partial_count(long long n);
// sc[k] = a.b    (partial must hold dot_partial_count(n) doubles, counter one zeroed unsigned)
void launch_dot(long long n, const double* a, const double* b, double* partial, unsigned* counter, double* sc_out, cudaStream_t st);
void launch_axpy(long long n, double* out, const double* x0, const double* p, double alpha, cudaStream_t st);  // out = x0 + alpha*p
// initX + xTilde (Optimizer.cpp:472-493, 585-610): x += (dt v + dt^2 g) on free verts
void launch_warm_start(int nV, double* x, const double* vel, const unsigned char* fixed, double dt, double gx, double gy, double gz,
                       cudaStream_t st);
void launch_xtilde(int nV, double* xt, const double* xn, const double* vel, const unsigned char* fixed, double dt, double gx, double gy,
                   double gz, cudaStream_t st);
void launch_velocity(int nV, double* vel, const double* x, const double* xn, double dt, cudaStream_t st);

// ---- fused L-BFGS iteration kernels (compact two-loop recursion: all inner products of an iteration are taken in two
// multi-dot passes against the Gram matrix of the history, DOTTimeStepper.cpp:389-398, 459-466 restated) ----
constexpr int LB_MAXH = 8;
// Speculative enqueue (stepper.cu): the kernels of iteration i+1 are launched before the host has seen the result of iteration i.
// They carry a pointer `go` to a device flag written by the LAST kernel of iteration i (1: the step was accepted, the new pair kept,
// not converged - exactly what the host assumed when it enqueued them); with *go == 0 every kernel returns at once.  go == nullptr: run.
struct HistList {            // active pairs, oldest -> newest, as buffer slots
    const int* go;
    int n;
    int slot[LB_MAXH];
    const double* S[LB_MAXH];  // by position
    const double* Y[LB_MAXH];
};
struct DotPairs {
    const int* go;
    int n;
    const double* a[12];
    const double* b[12];
    int out[12];             // index into the scalar array
};
// device scalar slots
enum ScalarSlot {
    SC_E = 0, SC_GG = 1, SC_PG = 2, SC_PHP = 3, SC_ALPHA = 4, SC_P0G = 5, SC_YS_NEW = 6, SC_DOT = 7,
    SC_SG = 8,    // s_i . g      by slot
    SC_YP = 16,   // y_i . p0     by slot
    SC_XI = 24,   // first-loop coefficients by slot
    SC_SY = 32,   // Gram matrix (s_i . y_j) at [SC_SY + 8*i + j], by slots
    SC_COUNT = 96,
    SC_EPREV = 7   // energy of the last accepted point (speculative iterations compare against it on the device); shares the slot of SC_DOT
};
int multidot_partial_count();  // sized for the widest multi-reduction (4 + 3 * LB_MAXH values)
int multidot_blocks(long long n);  // fixed grid of the deterministic multi-reduction kernels
// sc[P.out[j]] = a_j . b_j for all pairs in one pass (deterministic: fixed grid, block partials, last block adds them in order)
void launch_dots(long long n, const DotPairs& P, double* partial, unsigned* counter, double* sc, cudaStream_t st);
// q = -g - sum_i xi_i y_i with xi from the compact first loop (needs sc[SC_SG+slot], sc[SC_SY..]); stores xi to sc[SC_XI+slot]
void launch_lbfgs_q(long long n, double* q, const double* g, const HistList& H, double* sc, cudaStream_t st);
// p = p0 + sum_i (xi_i - beta_i) s_i with beta from the compact second loop (needs sc[SC_YP+slot], sc[SC_P0G]); stores p.g to sc[SC_PG]
void launch_lbfgs_p(long long n, double* p, const HistList& H, double* sc, cudaStream_t st);
// sc[SC_PHP] = p^T A p and sc[SC_ALPHA] = clamp(-sc[SC_PG] / sc[SC_PHP], 0.1, 1)   (Optimizer::initStepSize, Optimizer.cpp:1076-1093)
void launch_quadform_alpha(int n, const int* ia, const int* ja, const double* a, const double* p, double* partial, unsigned* counter,
                           double* sc, cudaStream_t st, const int* go = nullptr);
// out = x0 + alpha p with alpha read from the device (alpha_dev) or given (alpha_dev == nullptr)
void launch_axpy_dev(long long n, double* out, const double* x0, const double* p, const double* alpha_dev, double alpha_host,
                     cudaStream_t st, const int* go = nullptr);
// new pair s = alpha p, y = g_new - g_old written to S_new / Y_new, and in the same pass: sc[SC_GG] = g_new.g_new,
// sc[SC_SY + 8*sl + sl] = y.s, sc[SC_SY + 8*slot_i + sl] = s_i.y, sc[SC_SY + 8*sl + slot_i] = s.y_i
void launch_pair_dots(long long n, const double* p, const double* g_new, const double* g_old, double* S_new, double* Y_new, int sl,
                      const double* alpha_dev, double alpha_host, const HistList& H, double* partial, unsigned* counter, double* sc,
                      cudaStream_t st, const PeerSrc* src = nullptr, double* g_new_out = nullptr, bool with_energy = false);
// (src != nullptr: g_new is still one slot per rank in peer-written memory; the kernel adds the slots in rank order, writes the
//  reduced gradient to g_new_out and, with_energy, the reduced energy riding at index n to sc[SC_E])

// scatter/average of the subdomain solutions fused with the inner products p . P.a[j] -> sc[P.out[j]] (P.b is ignored)
void launch_scatter_avg_dots(int ndof, const int* cptr, const int* cidx, const double* xs, const int* dup, double* p, const DotPairs& P,
                             double* partial, unsigned* counter, double* sc, cudaStream_t st);

// multi-GPU: p = all-reduced sum of the subdomain solutions -> divide by dup and take p . P.a[j] in one pass
void launch_divdup_dots(int ndof, const int* dup, double* p, const DotPairs& P, double* partial, unsigned* counter, double* sc, cudaStream_t st,
                        const PeerSrc* src = nullptr);
// multi-GPU: this rank's sum of its subdomain copies -> its slot in every rank's peer buffer + flag (first half of the all-reduce of p)
void launch_scatter_push(int ndof, const int* cptr, const int* cidx, const double* xs, const PeerDst& D, cudaStream_t st);

// ---- preconditioner gather / scatter (DOTTimeStepper.cpp:414-450) ----
// p[d] = (sum over the subdomain copies of dof d, in subdomain order) / dup
void launch_scatter_avg(int ndof, const int* cptr, const int* cidx, const double* xs, const int* dup, double* p, cudaStream_t st);

}  // namespace dotgpu
