// wire_formats.cpp -- see wire_formats.hpp.
#include "wire_formats.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace lvreg_host {

namespace {

// XCDR1 little-endian writer: primitives are aligned to their size relative to the start of the body (after the
// 4-byte encapsulation header); a sequence is a uint32 count followed by its elements; a string is a uint32 length
// that includes the terminating NUL, the bytes, the NUL
struct CdrWriter {
    std::vector<uint8_t> buf;
    CdrWriter() { buf = {0x00, 0x01, 0x00, 0x00}; }                      // CDR_LE, no options
    void align(size_t a) {
        const size_t body = buf.size() - 4;
        const size_t pad = (a - body % a) % a;
        buf.insert(buf.end(), pad, 0);
    }
    template <class T> void put(T v) {
        align(sizeof(T));
        const uint8_t* p = reinterpret_cast<const uint8_t*>(&v);
        buf.insert(buf.end(), p, p + sizeof(T));
    }
    void put_string(const std::string& s) {
        put<uint32_t>((uint32_t)s.size() + 1);
        buf.insert(buf.end(), s.begin(), s.end());
        buf.push_back(0);
    }
    template <class T> void put_seq(const std::vector<T>& v) {
        put<uint32_t>((uint32_t)v.size());
        if (v.empty()) return;
        align(sizeof(T));
        const uint8_t* p = reinterpret_cast<const uint8_t*>(v.data());
        buf.insert(buf.end(), p, p + v.size() * sizeof(T));
    }
    void put_bytes(const void* data, size_t n) {
        put<uint32_t>((uint32_t)n);
        const uint8_t* p = reinterpret_cast<const uint8_t*>(data);
        buf.insert(buf.end(), p, p + n);
    }
};

struct CdrReader {
    const uint8_t* d;
    size_t n, pos = 4;
    CdrReader(const uint8_t* data, size_t size) : d(data), n(size) {
        if (size < 4 || data[0] != 0x00 || data[1] != 0x01) throw std::runtime_error("CloudInfo: not a little-endian CDR body");
    }
    void need(size_t k) const { if (pos + k > n) throw std::runtime_error("CloudInfo: truncated CDR body"); }
    void align(size_t a) { pos += (a - (pos - 4) % a) % a; }
    template <class T> T get() {
        align(sizeof(T));
        need(sizeof(T));
        T v;
        std::memcpy(&v, d + pos, sizeof(T));
        pos += sizeof(T);
        return v;
    }
    std::string get_string() {
        const uint32_t len = get<uint32_t>();
        need(len);
        std::string s(reinterpret_cast<const char*>(d + pos), len ? len - 1 : 0);
        pos += len;
        return s;
    }
    template <class T> std::vector<T> get_seq() {
        const uint32_t cnt = get<uint32_t>();
        std::vector<T> v(cnt);
        if (cnt) {
            align(sizeof(T));
            need((size_t)cnt * sizeof(T));
            std::memcpy(v.data(), d + pos, (size_t)cnt * sizeof(T));
            pos += (size_t)cnt * sizeof(T);
        }
        return v;
    }
};

void put_header(CdrWriter& w, double stamp, const std::string& frame_id) {
    const double sec = std::floor(stamp);
    w.put<int32_t>((int32_t)sec);                                        // builtin_interfaces/Time
    w.put<uint32_t>((uint32_t)std::llround((stamp - sec) * 1e9));
    w.put_string(frame_id);
}
double get_header(CdrReader& r, std::string* frame_id) {
    const int32_t sec = r.get<int32_t>();
    const uint32_t nsec = r.get<uint32_t>();
    const std::string f = r.get_string();
    if (frame_id) *frame_id = f;
    return (double)sec + 1e-9 * (double)nsec;
}

// sensor_msgs/PointCloud2 of a pcl::PointXYZI cloud, as pcl::toROSMsg fills it
void put_cloud(CdrWriter& w, const Cloud& c, double stamp, const std::string& frame_id) {
    put_header(w, stamp, frame_id);
    w.put<uint32_t>(1);                                                  // height
    w.put<uint32_t>((uint32_t)c.size());                                 // width
    static const char* names[4] = {"x", "y", "z", "intensity"};
    static const uint32_t offs[4] = {0, 4, 8, 16};
    w.put<uint32_t>(4);                                                  // fields[]
    for (int i = 0; i < 4; ++i) {
        w.put_string(names[i]);
        w.put<uint32_t>(offs[i]);
        w.put<uint8_t>(7);                                               // FLOAT32
        w.put<uint32_t>(1);                                              // count
    }
    w.put<uint8_t>(0);                                                   // is_bigendian
    w.put<uint32_t>((uint32_t)sizeof(PointType));                        // point_step
    w.put<uint32_t>((uint32_t)(c.size() * sizeof(PointType)));           // row_step
    w.put_bytes(c.data(), c.size() * sizeof(PointType));                 // data
    w.put<uint8_t>(1);                                                   // is_dense
}
Cloud get_cloud(CdrReader& r) {
    get_header(r, nullptr);
    const uint32_t height = r.get<uint32_t>(), width = r.get<uint32_t>();
    const uint32_t nf = r.get<uint32_t>();
    uint32_t off_x = 0, off_y = 4, off_z = 8, off_i = 16;
    for (uint32_t i = 0; i < nf; ++i) {
        const std::string name = r.get_string();
        const uint32_t off = r.get<uint32_t>();
        const uint8_t type = r.get<uint8_t>();
        r.get<uint32_t>();
        if (type != 7) continue;
        if (name == "x") off_x = off; else if (name == "y") off_y = off; else if (name == "z") off_z = off;
        else if (name == "intensity") off_i = off;
    }
    if (r.get<uint8_t>() != 0) throw std::runtime_error("CloudInfo: big-endian point data is not supported");
    const uint32_t step = r.get<uint32_t>();
    r.get<uint32_t>();                                                   // row_step
    const std::vector<uint8_t> data = r.get_seq<uint8_t>();
    r.get<uint8_t>();                                                    // is_dense
    const size_t n = (size_t)height * width;
    if (step < 16 || data.size() < n * step) throw std::runtime_error("CloudInfo: point data shorter than width x height");
    Cloud c(n);
    for (size_t i = 0; i < n; ++i) {
        const uint8_t* p = data.data() + i * step;
        float x, y, z, it = 0.f;
        std::memcpy(&x, p + off_x, 4); std::memcpy(&y, p + off_y, 4); std::memcpy(&z, p + off_z, 4);
        if (off_i + 4 <= step) std::memcpy(&it, p + off_i, 4);
        c[i] = make_point(x, y, z, it);
    }
    return c;
}

}  // namespace

std::vector<uint8_t> serialize_cloud_info(const CloudInfo& m, const std::string& frame_id) {
    CdrWriter w;
    put_header(w, m.stamp, frame_id);
    w.put_seq(m.start_ring_index);
    w.put_seq(m.end_ring_index);
    w.put_seq(m.point_col_ind);
    w.put_seq(m.point_range);
    w.put<int64_t>(m.imu_available);
    w.put<int64_t>(m.odom_available);
    w.put<float>(m.imu_roll_init); w.put<float>(m.imu_pitch_init); w.put<float>(m.imu_yaw_init);
    w.put<float>(m.initial_guess_x); w.put<float>(m.initial_guess_y); w.put<float>(m.initial_guess_z);
    w.put<float>(m.initial_guess_roll); w.put<float>(m.initial_guess_pitch); w.put<float>(m.initial_guess_yaw);
    w.put<int64_t>(m.odom_reset_id);
    put_cloud(w, m.cloud_deskewed, m.stamp, frame_id);
    put_cloud(w, m.cloud_corner, m.stamp, frame_id);
    put_cloud(w, m.cloud_surface, m.stamp, frame_id);
    return std::move(w.buf);
}

CloudInfo deserialize_cloud_info(const uint8_t* data, size_t size, std::string* frame_id) {
    CdrReader r(data, size);
    CloudInfo m;
    m.stamp = get_header(r, frame_id);
    m.start_ring_index = r.get_seq<int32_t>();
    m.end_ring_index = r.get_seq<int32_t>();
    m.point_col_ind = r.get_seq<int32_t>();
    m.point_range = r.get_seq<float>();
    m.imu_available = r.get<int64_t>();
    m.odom_available = r.get<int64_t>();
    m.imu_roll_init = r.get<float>(); m.imu_pitch_init = r.get<float>(); m.imu_yaw_init = r.get<float>();
    m.initial_guess_x = r.get<float>(); m.initial_guess_y = r.get<float>(); m.initial_guess_z = r.get<float>();
    m.initial_guess_roll = r.get<float>(); m.initial_guess_pitch = r.get<float>(); m.initial_guess_yaw = r.get<float>();
    m.odom_reset_id = r.get<int64_t>();
    m.cloud_deskewed = get_cloud(r);
    m.cloud_corner = get_cloud(r);
    m.cloud_surface = get_cloud(r);
    return m;
}

// ---- PCD ----
namespace {
bool write_pcd(const std::string& path, const char* fields, const char* sizes, const char* types, const char* counts,
               size_t n, const std::vector<uint8_t>& body) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS %s\nSIZE %s\nTYPE %s\nCOUNT %s\n"
                    "WIDTH %zu\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA binary\n",
                 fields, sizes, types, counts, n, n);
    const bool ok = body.empty() || std::fwrite(body.data(), 1, body.size(), f) == body.size();
    return std::fclose(f) == 0 && ok;
}
}  // namespace

bool save_pcd_binary(const std::string& path, const Cloud& cloud) {
    std::vector<uint8_t> body(cloud.size() * 16);
    for (size_t i = 0; i < cloud.size(); ++i) {
        const float v[4] = {cloud[i].x, cloud[i].y, cloud[i].z, cloud[i].intensity};
        std::memcpy(body.data() + i * 16, v, 16);
    }
    return write_pcd(path, "x y z intensity", "4 4 4 4", "F F F F", "1 1 1 1", cloud.size(), body);
}

bool save_pcd_binary(const std::string& path, const std::vector<PointTypePose>& poses) {
    std::vector<uint8_t> body(poses.size() * 36);
    for (size_t i = 0; i < poses.size(); ++i) {
        const PointTypePose& p = poses[i];
        const float v[7] = {p.x, p.y, p.z, p.intensity, p.roll, p.pitch, p.yaw};
        std::memcpy(body.data() + i * 36, v, 28);
        std::memcpy(body.data() + i * 36 + 28, &p.time, 8);
    }
    return write_pcd(path, "x y z intensity roll pitch yaw time", "4 4 4 4 4 4 4 8", "F F F F F F F F", "1 1 1 1 1 1 1 1",
                     poses.size(), body);
}

bool load_pcd_binary_xyzi(const std::string& path, Cloud* cloud) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    char line[512];
    size_t points = 0, stride = 0;
    bool binary = false;
    while (std::fgets(line, sizeof(line), f)) {
        if (!std::strncmp(line, "SIZE", 4)) {
            stride = 0;
            for (char* tok = std::strtok(line + 4, " \n"); tok; tok = std::strtok(nullptr, " \n")) stride += (size_t)std::atoi(tok);
        } else if (!std::strncmp(line, "POINTS", 6)) {
            points = (size_t)std::strtoull(line + 6, nullptr, 10);
        } else if (!std::strncmp(line, "DATA", 4)) {
            binary = std::strstr(line, "binary") != nullptr && std::strstr(line, "compressed") == nullptr;
            break;
        }
    }
    bool ok = binary && stride >= 16;
    if (ok) {
        std::vector<uint8_t> body(points * stride);
        ok = body.empty() || std::fread(body.data(), 1, body.size(), f) == body.size();
        if (ok) {
            cloud->resize(points);
            for (size_t i = 0; i < points; ++i) {
                float v[4];
                std::memcpy(v, body.data() + i * stride, 16);
                (*cloud)[i] = make_point(v[0], v[1], v[2], v[3]);
            }
        }
    }
    std::fclose(f);
    return ok;
}

}  // namespace lvreg_host
