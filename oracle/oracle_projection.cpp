// oracle_projection.cpp -- CPU restatement of the LiDAR front end's per-point work:
//   IP:495-526  findRotation         IP:538-569  deskewPoint
//   IP:571-623  projectPointCloud    IP:625-647  cloudExtraction
// (IP: = /root/reference/lidar_odometry/src/imageProjection.cpp).  The IMU integration that fills
// imuTime / imuRot{X,Y,Z} (IP:340-408) is sequential host work and is an input here.
//
// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// PARITY PIN STATUS: UNPINNED by the reference (no fixtures; PCL / Eigen absent).  Restated from
// Eigen 3.4: Affine3f::inverse() (3x3 cofactor inverse, translation = -inv * t) and the
// Affine3f * Affine3f product, pcl::getTransformation.  One deliberate deviation: the per-point
// sin / cos of getTransformation and the atan2 of the Velodyne / Ouster column index are evaluated
// in double and rounded to float (here and on the device) instead of through glibc's sinf / cosf /
// atan2f, which are accurate to <1 ulp but not correctly rounded -- a device implementation cannot
// reproduce their last bit for every one of the ~10^5 angles of a scan, the correctly rounded value
// it can.  The difference to the reference is at most one ulp of a rotation-matrix entry.
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace {

// Eigen::internal::compute_inverse<Matrix3f> (cofactors, determinant along column 0)
void inverse3(const float m[3][3], float r[3][3]) {
    auto cof = [&](int i, int j) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return m[i1][j1] * m[i2][j2] - m[i1][j2] * m[i2][j1];
    };
    const float c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
    const float det = (c0 * m[0][0] + c1 * m[1][0]) + c2 * m[2][0];
    const float invdet = 1.0f / det;
    r[0][0] = c0 * invdet; r[0][1] = c1 * invdet; r[0][2] = c2 * invdet;
    r[1][0] = cof(0, 1) * invdet; r[1][1] = cof(1, 1) * invdet; r[1][2] = cof(2, 1) * invdet;
    r[2][0] = cof(0, 2) * invdet; r[2][1] = cof(1, 2) * invdet; r[2][2] = cof(2, 2) * invdet;
}

// row-major 3x4 affine inverse / product, Eigen::Transform<float,3,Affine>
void affine_inverse(const float T[12], float R[12]) {
    float m[3][3], inv[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) m[i][j] = T[4 * i + j];
    inverse3(m, inv);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R[4 * i + j] = inv[i][j];
        R[4 * i + 3] = -((inv[i][0] * T[3] + inv[i][1] * T[7]) + inv[i][2] * T[11]);
    }
}
// pcl::getTransformation(0, 0, 0, roll, pitch, yaw), trig correctly rounded
void rotation_affine(const float rot[3], float T[12]) {
    const float roll = rot[0], pitch = rot[1], yaw = rot[2];
    const float A = (float)std::cos((double)yaw), B = (float)std::sin((double)yaw);
    const float C = (float)std::cos((double)pitch), D = (float)std::sin((double)pitch);
    const float E = (float)std::cos((double)roll), F = (float)std::sin((double)roll);
    const float DE = D * E, DF = D * F;
    T[0] = A * C;  T[1] = A * DF - B * E;  T[2]  = B * F + A * DE;  T[3]  = 0.f;
    T[4] = B * C;  T[5] = A * E + B * DF;  T[6]  = B * DE - A * F;  T[7]  = 0.f;
    T[8] = -D;     T[9] = C * F;           T[10] = C * E;           T[11] = 0.f;
}
void affine_mul(const float A[12], const float B[12], float C[12]) {
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j)
            C[4 * i + j] = (A[4 * i] * B[j] + A[4 * i + 1] * B[4 + j]) + A[4 * i + 2] * B[8 + j];
        C[4 * i + 3] = ((A[4 * i] * B[3] + A[4 * i + 1] * B[7]) + A[4 * i + 2] * B[11]) + A[4 * i + 3];
    }
}

}  // namespace

// findRotation, IP:495-526
extern "C" void orc_find_rotation(double point_time, const double* imu_time, const double* rx, const double* ry,
                                  const double* rz, int imu_pointer_cur, float rot[3]) {
    int f = 0;
    while (f < imu_pointer_cur) {
        if (point_time < imu_time[f]) break;
        ++f;
    }
    if (point_time > imu_time[f] || f == 0) {
        rot[0] = (float)rx[f]; rot[1] = (float)ry[f]; rot[2] = (float)rz[f];
    } else {
        const int b = f - 1;
        const double rf = (point_time - imu_time[b]) / (imu_time[f] - imu_time[b]);
        const double rb = (imu_time[f] - point_time) / (imu_time[f] - imu_time[b]);
        rot[0] = (float)(rx[f] * rf + rx[b] * rb);
        rot[1] = (float)(ry[f] * rf + ry[b] * rb);
        rot[2] = (float)(rz[f] * rf + rz[b] * rb);
    }
}

// projectPointCloud + cloudExtraction.  pts: n x 4 {x,y,z,intensity}; ring / rel_time: n entries.
// Outputs sized n_scan * horizon (cloud, range, col) and n_scan (start / end).  Returns the size of
// the extracted cloud.
extern "C" size_t orc_project_cloud(const float* pts, const uint16_t* ring, const float* rel_time, size_t n,
                                    const orc_projection_params* P, float* extracted, float* point_range,
                                    int32_t* point_col_ind, int32_t* start_ring_index, int32_t* end_ring_index) {
    const int NS = P->n_scan, H = P->horizon_scan;
    std::vector<float> rangeMat((size_t)NS * H, FLT_MAX);
    std::vector<float> full(4 * (size_t)NS * H, 0.f);
    std::vector<int> columnIdnCountVec(NS, 0);
    bool first = true;
    float start_inv[12] = {0};
    for (size_t i = 0; i < n; ++i) {
        const float* p = pts + 4 * i;
        const float range = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
        if (range < P->lidar_min_range || range > P->lidar_max_range) continue;
        const int row = ring[i];
        if (row < 0 || row >= NS) continue;
        if (row % P->downsample_rate != 0) continue;
        int col = -1;
        if (P->sensor == 0 || P->sensor == 1) {
            const float horizonAngle = (float)((float)((float)std::atan2((double)p[0], (double)p[1]) * 180) / M_PI);
            const float ang_res_x = 360.0 / float(H);
            col = (int)(-std::round((horizonAngle - 90.0) / ang_res_x) + H / 2);
            if (col >= H) col -= H;
        } else {
            col = columnIdnCountVec[row];
            columnIdnCountVec[row] += 1;
        }
        if (col < 0 || col >= H) continue;
        if (rangeMat[(size_t)row * H + col] != FLT_MAX) continue;
        float q[4] = {p[0], p[1], p[2], p[3]};
        if (P->deskew) {                                          // deskewPoint, IP:538-569
            float rot[3];
            orc_find_rotation(P->time_scan_cur + (double)rel_time[i], P->imu_time, P->imu_rot_x, P->imu_rot_y,
                              P->imu_rot_z, P->imu_pointer_cur, rot);
            float T[12];
            rotation_affine(rot, T);                               // findPosition returns zeros (IP:528-536)
            if (first) {
                affine_inverse(T, start_inv);
                first = false;
            }
            float Bt[12];
            affine_mul(start_inv, T, Bt);
            q[0] = Bt[0] * p[0] + Bt[1] * p[1] + Bt[2] * p[2] + Bt[3];
            q[1] = Bt[4] * p[0] + Bt[5] * p[1] + Bt[6] * p[2] + Bt[7];
            q[2] = Bt[8] * p[0] + Bt[9] * p[1] + Bt[10] * p[2] + Bt[11];
        }
        rangeMat[(size_t)row * H + col] = range;
        std::memcpy(&full[4 * ((size_t)row * H + col)], q, sizeof(q));
    }
    size_t count = 0;
    for (int i = 0; i < NS; ++i) {
        start_ring_index[i] = (int32_t)count - 1 + 5;
        for (int j = 0; j < H; ++j)
            if (rangeMat[(size_t)i * H + j] != FLT_MAX) {
                point_col_ind[count] = j;
                point_range[count] = rangeMat[(size_t)i * H + j];
                std::memcpy(extracted + 4 * count, &full[4 * ((size_t)i * H + j)], 4 * sizeof(float));
                ++count;
            }
        end_ring_index[i] = (int32_t)count - 1 - 5;
    }
    return count;
}
