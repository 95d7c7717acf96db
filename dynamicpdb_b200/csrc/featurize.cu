// Input featurisation of a trajectory window on the device.
//
// One thread per (frame, residue): from the atom37 coordinates of a window [nf,N,37,3] it produces what the reference's
// loader computes on the host, per sample, with a chain of torch gathers and a CPU eigen-decomposition
// (src/data/Dfold_data_loader_dynamic.py:192-259, :323-330):
//   rigids_0   backbone frame (Gram-Schmidt on C, CA, N; openfold/data/data_transforms.py:755-842 group 0, composed with
//              diag(-1, 1, -1)) as quaternion + translation; the quaternion comes from the closed form instead of
//              rot_to_quat's symmetric 4x4 eigh (openfold/utils/rigid_utils.py:208-227), whose sign is arbitrary anyway
//   torsions   sin / cos of pre-omega, phi, psi, chi1..4, the pi-periodic alternative and the mask
//              (data_transforms.py:923-1088), in fp64 as the reference (which warns that fp32 is too imprecise there)
// All 3x3 arithmetic stays in registers; the masked coordinates (Dfold_data_loader_dynamic.py:85) are formed on the fly.
#include "common.cuh"

namespace dfold {
namespace {

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 scale(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// Rigid.from_3_points (rigid_utils.py:1233-1275): columns e0, e1, e2
__device__ __forceinline__ void frame3(V3 p_neg_x, V3 origin, V3 p_xy, double eps, V3& e0, V3& e1, V3& e2) {
    e0 = sub(origin, p_neg_x);
    e1 = sub(p_xy, origin);
    e0 = scale(e0, 1.0 / sqrt(dot(e0, e0) + eps));
    e1 = sub(e1, scale(e0, dot(e0, e1)));
    e1 = scale(e1, 1.0 / sqrt(dot(e1, e1) + eps));
    e2 = cross(e0, e1);
}

struct FeatParams {
    const float* pos;        // [nf,N,37,3]
    const float* amask;      // [N,37]
    const long* aatype;      // [N]
    const long* chi_idx;     // [21,4,4]
    const float* chi_mask;   // [21,4]
    const float* chi_pi;     // [21,4]
    int nf, N;
    float* rigids;           // [nf,N,7]
    double* tor; double* alt; double* tmask;     // [nf,N,7,2] x2, [nf,N,7]
};

__global__ void featurize_kernel(const FeatParams p) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)p.nf * p.N) return;
    const int i = (int)(idx % p.N);
    const float* P = p.pos + idx * 111;
    const float* M = p.amask + (long)i * 37;
    const float* Pp = P - 111;                                // previous residue of the same frame
    const float* Mp = M - 37;
    const bool has_prev = i > 0;
    auto at = [&](const float* base, const float* m, int a) -> V3 {
        const double w = m[a];
        return {base[3 * a] * w, base[3 * a + 1] * w, base[3 * a + 2] * w};
    };
    const V3 zero = {0, 0, 0};
    // ---- backbone frame: from_3_points(C, CA, N) o diag(-1, 1, -1) ----
    {
        V3 e0, e1, e2;
        const V3 ca = at(P, M, 1);
        frame3(at(P, M, 2), ca, at(P, M, 0), 1e-8, e0, e1, e2);
        // rotation matrix columns (-e0, e1, -e2); closed-form quaternion (Shepperd: largest of w, x, y, z first)
        const double r00 = -e0.x, r10 = -e0.y, r20 = -e0.z, r01 = e1.x, r11 = e1.y, r21 = e1.z, r02 = -e2.x, r12 = -e2.y, r22 = -e2.z;
        double w, x, y, z;
        const double tr = r00 + r11 + r22;
        if (tr > 0) {
            const double s = sqrt(tr + 1.0) * 2; w = 0.25 * s; x = (r21 - r12) / s; y = (r02 - r20) / s; z = (r10 - r01) / s;
        } else if (r00 > r11 && r00 > r22) {
            const double s = sqrt(1.0 + r00 - r11 - r22) * 2; w = (r21 - r12) / s; x = 0.25 * s; y = (r01 + r10) / s; z = (r02 + r20) / s;
        } else if (r11 > r22) {
            const double s = sqrt(1.0 + r11 - r00 - r22) * 2; w = (r02 - r20) / s; x = (r01 + r10) / s; y = 0.25 * s; z = (r12 + r21) / s;
        } else {
            const double s = sqrt(1.0 + r22 - r00 - r11) * 2; w = (r10 - r01) / s; x = (r02 + r20) / s; y = (r12 + r21) / s; z = 0.25 * s;
        }
        float* o = p.rigids + idx * 7;
        o[0] = (float)w; o[1] = (float)x; o[2] = (float)y; o[3] = (float)z; o[4] = (float)ca.x; o[5] = (float)ca.y; o[6] = (float)ca.z;
    }
    // ---- torsions ----
    long aa = p.aatype[i];
    aa = aa > 20 ? 20 : aa;
    for (int t = 0; t < 7; ++t) {
        V3 a0, a1, a2, a3;
        double mk;
        if (t == 0) {            // pre-omega: prev CA, prev C, N, CA
            a0 = has_prev ? at(Pp, Mp, 1) : zero; a1 = has_prev ? at(Pp, Mp, 2) : zero; a2 = at(P, M, 0); a3 = at(P, M, 1);
            mk = has_prev ? (double)Mp[1] * Mp[2] * M[0] * M[1] : 0.0;
        } else if (t == 1) {     // phi: prev C, N, CA, C
            a0 = has_prev ? at(Pp, Mp, 2) : zero; a1 = at(P, M, 0); a2 = at(P, M, 1); a3 = at(P, M, 2);
            mk = has_prev ? (double)Mp[2] * M[0] * M[1] * M[2] : 0.0;
        } else if (t == 2) {     // psi: N, CA, C, O
            a0 = at(P, M, 0); a1 = at(P, M, 1); a2 = at(P, M, 2); a3 = at(P, M, 4);
            mk = (double)M[0] * M[1] * M[2] * M[4];
        } else {
            const long* ci = p.chi_idx + (aa * 4 + (t - 3)) * 4;
            a0 = at(P, M, (int)ci[0]); a1 = at(P, M, (int)ci[1]); a2 = at(P, M, (int)ci[2]); a3 = at(P, M, (int)ci[3]);
            mk = (double)p.chi_mask[aa * 4 + (t - 3)] * M[ci[0]] * M[ci[1]] * M[ci[2]] * M[ci[3]];
        }
        V3 e0, e1, e2;
        frame3(a1, a2, a0, 1e-8, e0, e1, e2);
        const V3 d = sub(a3, a2);
        const double ry = dot(e1, d), rz = dot(e2, d);
        const double den = sqrt(rz * rz + ry * ry + 1e-8);
        const double sgn = (t == 2) ? -1.0 : 1.0;
        const double s = rz / den * sgn, c = ry / den * sgn;
        const double mir = (t < 3) ? 1.0 : 1.0 - 2.0 * (double)p.chi_pi[aa * 4 + (t - 3)];
        const long o = idx * 7 + t;
        p.tor[2 * o] = s; p.tor[2 * o + 1] = c;
        p.alt[2 * o] = s * mir; p.alt[2 * o + 1] = c * mir;
        p.tmask[o] = mk;
    }
}

}  // namespace
}  // namespace dfold

using namespace dfold;

// pos [nf,N,37,3] fp32, atom_mask [N,37] fp32, aatype [N] int64, tables chi_idx [21,4,4] int64, chi_mask / chi_pi [21,4] fp32
// -> rigids_0 [nf,N,7] fp32 (quat wxyz + CA), torsion sin/cos, alternative, mask (fp64)
extern "C" int dfold_featurize_window(const float* pos, const float* atom_mask, const long* aatype, const long* chi_idx,
                                      const float* chi_mask, const float* chi_pi, int nf, int N, float* rigids, double* tor,
                                      double* alt, double* tmask, void* stream) {
    DFOLD_REQUIRE(nf > 0 && N > 0, "dfold_featurize_window: empty window");
    FeatParams p;
    p.pos = pos; p.amask = atom_mask; p.aatype = aatype; p.chi_idx = chi_idx; p.chi_mask = chi_mask; p.chi_pi = chi_pi;
    p.nf = nf; p.N = N; p.rigids = rigids; p.tor = tor; p.alt = alt; p.tmask = tmask;
    const long n = (long)nf * N;
    featurize_kernel<<<(unsigned)cdiv(n, 128), 128, 0, as_stream(stream)>>>(p);
    return check_launch("featurize_kernel");
}
