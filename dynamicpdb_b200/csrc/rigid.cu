// Quaternion / rigid-frame kernels: all 3x3 arithmetic stays in registers.
// Replaces (reference openfold/utils/rigid_utils.py): quat_to_rot :185-205, Rigid.apply :1104-1116,
// Rigid.invert_apply :1118-1130, Rigid.compose_q_update_vec :1039-1063 (+ Rotation.compose_q_update_vec :587-616,
// quat_multiply_by_vec :266-275) and their autograd.
#include "common.cuh"

namespace dfold {
namespace {

__global__ void quat_to_rot_fwd_kernel(const float* __restrict__ q, float* __restrict__ R, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 qq = *reinterpret_cast<const float4*>(q + 4 * i);
    float r[9];
    quat_to_rot9(qq.x, qq.y, qq.z, qq.w, r);
#pragma unroll
    for (int k = 0; k < 9; ++k) R[9 * i + k] = r[k];
}

__global__ void quat_to_rot_bwd_kernel(const float* __restrict__ q, const float* __restrict__ dR, float* __restrict__ dq, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 qq = *reinterpret_cast<const float4*>(q + 4 * i);
    float g[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) g[k] = dR[9 * i + k];
    float dw, dx, dy, dz;
    quat_to_rot9_bwd(qq.x, qq.y, qq.z, qq.w, g, dw, dx, dy, dz);
    *reinterpret_cast<float4*>(dq + 4 * i) = make_float4(dw, dx, dy, dz);
}

// frames n = F*N (index f*N+i); m points per frame.  pts row base = f*pts_fs + i*m*3 (pts_fs = 0: shared by all f)
// out[n, m, 3] = R p + t            (inverse: R^T (p - t))
__global__ void rigid_apply_fwd_kernel(const float* __restrict__ quat, const float* __restrict__ trans,
                                       const float* __restrict__ pts, long pts_fs, float* __restrict__ out,
                                       long F, long N, int m, int inverse) {
    const long frame = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (frame >= F * N) return;
    const int lane = threadIdx.x & 31;
    const long f = frame / N, i = frame % N;
    const float4 q = *reinterpret_cast<const float4*>(quat + 4 * frame);
    float R[9];
    quat_to_rot9(q.x, q.y, q.z, q.w, R);
    const float tx = trans[3 * frame], ty = trans[3 * frame + 1], tz = trans[3 * frame + 2];
    const float* pp = pts + f * pts_fs + i * (long)m * 3;
    float* oo = out + frame * (long)m * 3;
    for (int k = lane; k < m; k += 32) {
        float x = pp[3 * k], y = pp[3 * k + 1], z = pp[3 * k + 2];
        float ox, oy, oz;
        if (!inverse) {
            ox = R[0] * x + R[1] * y + R[2] * z + tx;
            oy = R[3] * x + R[4] * y + R[5] * z + ty;
            oz = R[6] * x + R[7] * y + R[8] * z + tz;
        } else {
            x -= tx; y -= ty; z -= tz;
            ox = R[0] * x + R[3] * y + R[6] * z;
            oy = R[1] * x + R[4] * y + R[7] * z;
            oz = R[2] * x + R[5] * y + R[8] * z;
        }
        oo[3 * k] = ox; oo[3 * k + 1] = oy; oo[3 * k + 2] = oz;
    }
}

// dpts[n, m, 3] (dense per frame), dquat[n,4], dtrans[n,3]
__global__ void rigid_apply_bwd_kernel(const float* __restrict__ quat, const float* __restrict__ trans,
                                       const float* __restrict__ pts, long pts_fs, const float* __restrict__ dout,
                                       float* __restrict__ dpts, float* __restrict__ dquat, float* __restrict__ dtrans,
                                       long F, long N, int m, int inverse) {
    const long frame = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (frame >= F * N) return;
    const int lane = threadIdx.x & 31;
    const long f = frame / N, i = frame % N;
    const float4 q = *reinterpret_cast<const float4*>(quat + 4 * frame);
    float R[9];
    quat_to_rot9(q.x, q.y, q.z, q.w, R);
    const float tx = trans[3 * frame], ty = trans[3 * frame + 1], tz = trans[3 * frame + 2];
    const float* pp = pts + f * pts_fs + i * (long)m * 3;
    const float* gg = dout + frame * (long)m * 3;
    float* dp = dpts + frame * (long)m * 3;
    float dR[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float dt[3] = {0.f, 0.f, 0.f};
    for (int k = lane; k < m; k += 32) {
        const float gx = gg[3 * k], gy = gg[3 * k + 1], gz = gg[3 * k + 2];
        float x = pp[3 * k], y = pp[3 * k + 1], z = pp[3 * k + 2];
        if (!inverse) {
            // out_a = sum_b R[a][b] p_b + t_a
            dp[3 * k]     = R[0] * gx + R[3] * gy + R[6] * gz;
            dp[3 * k + 1] = R[1] * gx + R[4] * gy + R[7] * gz;
            dp[3 * k + 2] = R[2] * gx + R[5] * gy + R[8] * gz;
            dR[0] += gx * x; dR[1] += gx * y; dR[2] += gx * z;
            dR[3] += gy * x; dR[4] += gy * y; dR[5] += gy * z;
            dR[6] += gz * x; dR[7] += gz * y; dR[8] += gz * z;
            dt[0] += gx; dt[1] += gy; dt[2] += gz;
        } else {
            // out_a = sum_b R[b][a] (p - t)_b
            x -= tx; y -= ty; z -= tz;
            const float rx = R[0] * gx + R[1] * gy + R[2] * gz;
            const float ry = R[3] * gx + R[4] * gy + R[5] * gz;
            const float rz = R[6] * gx + R[7] * gy + R[8] * gz;
            dp[3 * k] = rx; dp[3 * k + 1] = ry; dp[3 * k + 2] = rz;
            dR[0] += x * gx; dR[1] += x * gy; dR[2] += x * gz;
            dR[3] += y * gx; dR[4] += y * gy; dR[5] += y * gz;
            dR[6] += z * gx; dR[7] += z * gy; dR[8] += z * gz;
            dt[0] -= rx; dt[1] -= ry; dt[2] -= rz;
        }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) dR[k] = warp_sum(dR[k]);
#pragma unroll
    for (int k = 0; k < 3; ++k) dt[k] = warp_sum(dt[k]);
    if (lane == 0) {
        float dw, dx, dy, dz;
        quat_to_rot9_bwd(q.x, q.y, q.z, q.w, dR, dw, dx, dy, dz);
        *reinterpret_cast<float4*>(dquat + 4 * frame) = make_float4(dw, dx, dy, dz);
        dtrans[3 * frame] = dt[0]; dtrans[3 * frame + 1] = dt[1]; dtrans[3 * frame + 2] = dt[2];
    }
}

// q' = normalise(q + m * q*(0,u)),  t' = t + m * R(q) v      upd = (u, v), mask nullable
__global__ void compose_fwd_kernel(const float* __restrict__ quat, const float* __restrict__ trans, const float* __restrict__ upd,
                                   const float* __restrict__ mask, float* __restrict__ qo, float* __restrict__ to, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 q = *reinterpret_cast<const float4*>(quat + 4 * i);
    const float a = q.x, b = q.y, c = q.z, d = q.w;
    const float x = upd[6 * i], y = upd[6 * i + 1], z = upd[6 * i + 2];
    const float vx = upd[6 * i + 3], vy = upd[6 * i + 4], vz = upd[6 * i + 5];
    const float m = mask ? mask[i] : 1.f;
    float nw = a + m * (-b * x - c * y - d * z);
    float nx = b + m * (a * x + c * z - d * y);
    float ny = c + m * (a * y - b * z + d * x);
    float nz = d + m * (a * z + b * y - c * x);
    const float inv = 1.f / sqrtf(nw * nw + nx * nx + ny * ny + nz * nz);
    *reinterpret_cast<float4*>(qo + 4 * i) = make_float4(nw * inv, nx * inv, ny * inv, nz * inv);
    float R[9];
    quat_to_rot9(a, b, c, d, R);
    to[3 * i]     = trans[3 * i]     + m * (R[0] * vx + R[1] * vy + R[2] * vz);
    to[3 * i + 1] = trans[3 * i + 1] + m * (R[3] * vx + R[4] * vy + R[5] * vz);
    to[3 * i + 2] = trans[3 * i + 2] + m * (R[6] * vx + R[7] * vy + R[8] * vz);
}

__global__ void compose_bwd_kernel(const float* __restrict__ quat, const float* __restrict__ upd, const float* __restrict__ mask,
                                   const float* __restrict__ dqo, const float* __restrict__ dto,
                                   float* __restrict__ dquat, float* __restrict__ dtrans, float* __restrict__ dupd, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 q = *reinterpret_cast<const float4*>(quat + 4 * i);
    const float a = q.x, b = q.y, c = q.z, d = q.w;
    const float x = upd[6 * i], y = upd[6 * i + 1], z = upd[6 * i + 2];
    const float vx = upd[6 * i + 3], vy = upd[6 * i + 4], vz = upd[6 * i + 5];
    const float m = mask ? mask[i] : 1.f;
    // recompute the un-normalised quaternion
    const float nw = a + m * (-b * x - c * y - d * z);
    const float nx = b + m * (a * x + c * z - d * y);
    const float ny = c + m * (a * y - b * z + d * x);
    const float nz = d + m * (a * z + b * y - c * x);
    const float inv = 1.f / sqrtf(nw * nw + nx * nx + ny * ny + nz * nz);
    const float ow = nw * inv, ox = nx * inv, oy = ny * inv, oz = nz * inv;
    const float4 g = *reinterpret_cast<const float4*>(dqo + 4 * i);
    const float dot = ow * g.x + ox * g.y + oy * g.z + oz * g.w;
    // d(un-normalised)
    const float gw = (g.x - ow * dot) * inv, gx = (g.y - ox * dot) * inv, gy = (g.z - oy * dot) * inv, gz = (g.w - oz * dot) * inv;
    const float hw = m * gw, hx = m * gx, hy = m * gy, hz = m * gz;     // gradient of the q*(0,u) product
    float da = gw + x * hx + y * hy + z * hz;
    float db = gx - x * hw - z * hy + y * hz;
    float dc = gy - y * hw + z * hx - x * hz;
    float dd = gz - z * hw - y * hx + x * hy;
    const float dux = -b * hw + a * hx + d * hy - c * hz;
    const float duy = -c * hw - d * hx + a * hy + b * hz;
    const float duz = -d * hw + c * hx - b * hy + a * hz;
    // translation part
    const float tx = dto[3 * i], ty = dto[3 * i + 1], tz = dto[3 * i + 2];
    float R[9];
    quat_to_rot9(a, b, c, d, R);
    const float dvx = m * (R[0] * tx + R[3] * ty + R[6] * tz);
    const float dvy = m * (R[1] * tx + R[4] * ty + R[7] * tz);
    const float dvz = m * (R[2] * tx + R[5] * ty + R[8] * tz);
    float dR[9] = {m * tx * vx, m * tx * vy, m * tx * vz, m * ty * vx, m * ty * vy, m * ty * vz, m * tz * vx, m * tz * vy, m * tz * vz};
    float ew, ex, ey, ez;
    quat_to_rot9_bwd(a, b, c, d, dR, ew, ex, ey, ez);
    *reinterpret_cast<float4*>(dquat + 4 * i) = make_float4(da + ew, db + ex, dc + ey, dd + ez);
    dtrans[3 * i] = tx; dtrans[3 * i + 1] = ty; dtrans[3 * i + 2] = tz;
    dupd[6 * i] = dux; dupd[6 * i + 1] = duy; dupd[6 * i + 2] = duz;
    dupd[6 * i + 3] = dvx; dupd[6 * i + 4] = dvy; dupd[6 * i + 5] = dvz;
}

}  // namespace
}  // namespace dfold

using namespace dfold;

extern "C" int dfold_quat_to_rot_fwd(const float* quat, float* rot, long n, void* stream) {
    DFOLD_REQUIRE(n > 0, "dfold_quat_to_rot_fwd: empty input");
    quat_to_rot_fwd_kernel<<<(unsigned)cdiv(n, 256), 256, 0, as_stream(stream)>>>(quat, rot, n);
    return check_launch("quat_to_rot_fwd_kernel");
}
extern "C" int dfold_quat_to_rot_bwd(const float* quat, const float* drot, float* dquat, long n, void* stream) {
    DFOLD_REQUIRE(n > 0, "dfold_quat_to_rot_bwd: empty input");
    quat_to_rot_bwd_kernel<<<(unsigned)cdiv(n, 256), 256, 0, as_stream(stream)>>>(quat, drot, dquat, n);
    return check_launch("quat_to_rot_bwd_kernel");
}
extern "C" int dfold_rigid_apply_fwd(const float* quat, const float* trans, const float* pts, long pts_fstride, float* out,
                                     long F, long N, int m, int inverse, void* stream) {
    DFOLD_REQUIRE(F > 0 && N > 0 && m > 0, "dfold_rigid_apply_fwd: empty input");
    rigid_apply_fwd_kernel<<<(unsigned)cdiv(F * N, 8), 256, 0, as_stream(stream)>>>(quat, trans, pts, pts_fstride, out, F, N, m, inverse);
    return check_launch("rigid_apply_fwd_kernel");
}
extern "C" int dfold_rigid_apply_bwd(const float* quat, const float* trans, const float* pts, long pts_fstride, const float* dout,
                                     float* dpts, float* dquat, float* dtrans, long F, long N, int m, int inverse, void* stream) {
    DFOLD_REQUIRE(F > 0 && N > 0 && m > 0, "dfold_rigid_apply_bwd: empty input");
    rigid_apply_bwd_kernel<<<(unsigned)cdiv(F * N, 8), 256, 0, as_stream(stream)>>>(quat, trans, pts, pts_fstride, dout, dpts, dquat, dtrans, F, N, m, inverse);
    return check_launch("rigid_apply_bwd_kernel");
}
extern "C" int dfold_compose_q_update_fwd(const float* quat, const float* trans, const float* upd6, const float* mask,
                                          float* quat_out, float* trans_out, long n, void* stream) {
    DFOLD_REQUIRE(n > 0, "dfold_compose_q_update_fwd: empty input");
    compose_fwd_kernel<<<(unsigned)cdiv(n, 256), 256, 0, as_stream(stream)>>>(quat, trans, upd6, mask, quat_out, trans_out, n);
    return check_launch("compose_fwd_kernel");
}
extern "C" int dfold_compose_q_update_bwd(const float* quat, const float* upd6, const float* mask, const float* dquat_out,
                                          const float* dtrans_out, float* dquat, float* dtrans, float* dupd6, long n, void* stream) {
    DFOLD_REQUIRE(n > 0, "dfold_compose_q_update_bwd: empty input");
    compose_bwd_kernel<<<(unsigned)cdiv(n, 256), 256, 0, as_stream(stream)>>>(quat, upd6, mask, dquat_out, dtrans_out, dquat, dtrans, dupd6, n);
    return check_launch("compose_bwd_kernel");
}
