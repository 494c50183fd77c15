// bos_synth.cpp -- synthetic bearing-only worlds for the benchmark configurations (host code, not timed; own library
// libbos_synth.so, see bos_synth.h).
//
// Mirrors the statistics of the reference's bundled dataset (data/slam2D_bearing_only_*.g2o): unit-step
// trajectory with half-turns at the row ends, landmarks uniform in the world, a range-limited sensor with
// a +-pi/2 field of view, bearing noise 3e-3 rad with omega = 1, odometry noise 1/sqrt(500) m and
// 1/sqrt(5000) rad with Omega = diag(500, 500, 5000), contiguous pose ids from 1200, sparse landmark ids.
// Counter-based RNG (splitmix64 + Box-Muller) so the same spec gives the same world everywhere.
// All values are rounded to float and widened, as utils/g2o_utils.cpp does with std::stof.
//
// Deviation from a literal "dead-reckoned initial guess": integrating noisy odometry over 10^4..10^5 poses
// drifts by kilometres, which no Gauss-Newton run recovers from; the initial guess here is the ground
// truth plus a smooth low-frequency drift and white noise, and landmarks are triangulated from it.
#include "bos_synth.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ULL;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}
struct Rng {
    uint64_t seed;
    double uniform(uint64_t stream, uint64_t idx) const {
        uint64_t h = splitmix64(seed ^ splitmix64(stream * 0x100000001b3ULL + 0x7f4a7c15ULL) ^ splitmix64(idx + 0x51ed270b7ULL));
        return ((h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    }
    double normal(uint64_t stream, uint64_t idx) const {
        double u1 = uniform(stream * 2 + 1000, idx), u2 = uniform(stream * 2 + 1001, idx);
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};
inline double f32(double v) { return (double)(float)v; }
inline double wrap(double a) {
    while (a < -3.141592653589793) a += 6.283185307179586;
    while (a >= 3.141592653589793) a -= 6.283185307179586;
    return a;
}

}  // namespace

struct bos_synth {
    bos_synth_spec spec;
    std::vector<int32_t> pose_ids, lm_ids, b_pose_id, b_lm_id, o_src_id, o_dst_id;
    std::vector<double> poses_true, poses_init, lms_true, b_z, o_z, o_omega;
};

extern "C" {

void bos_synth_default_spec(bos_synth_spec* s) {
    if (!s) return;
    std::memset(s, 0, sizeof(*s));
    s->n_poses = 1000; s->n_landmarks = 200; s->target_bearing_edges = 10000;
    s->seed = 0xB0500000ULL;
    s->bearing_sigma = 3e-3;
    s->odom_sigma_xy = 1.0 / std::sqrt(500.0);
    s->odom_sigma_theta = 1.0 / std::sqrt(5000.0);
    s->init_drift = 0.3;
    s->init_noise = 0.02;
}

int bos_synth_create(const bos_synth_spec* spec, bos_synth** out) {
    if (!spec || !out || spec->n_poses < 2 || spec->n_landmarks < 1 || spec->target_bearing_edges < 1) return 1;
    bos_synth* W = new bos_synth();
    W->spec = *spec;
    const int NP = spec->n_poses, NL = spec->n_landmarks;
    const Rng rng{spec->seed};
    const int cols = (int)std::ceil(std::sqrt((double)NP));
    const int rows = (NP + cols - 1) / cols;
    // ---- ground-truth trajectory: serpentine rows, heading 0 on even rows and pi on odd rows ------------
    std::vector<double> px(NP), py(NP), pth(NP);
    for (int i = 0; i < NP; i++) {
        const int r = i / cols, k = i % cols;
        px[i] = (r % 2 == 0) ? k : cols - 1 - k;
        py[i] = r;
        pth[i] = (r % 2 == 0) ? 0.0 : 3.141592653589793;
        // a gentle heading wobble so the field of view sweeps
        pth[i] = wrap(pth[i] + 0.25 * std::sin(0.37 * i));
    }
    // ---- landmarks -----------------------------------------------------------------------------------------
    const double x0 = -2.0, x1 = cols + 1.0, y0 = -2.0, y1 = rows + 1.0;
    const double area = (x1 - x0) * (y1 - y0);
    const double density = NL / area;
    double range = std::sqrt(2.0 * (double)spec->target_bearing_edges / (3.141592653589793 * NP * density));
    if (range < 1.5) range = 1.5;
    std::vector<double> lx(NL), ly(NL);
    for (int j = 0; j < NL; j++) {
        lx[j] = x0 + (x1 - x0) * rng.uniform(1, j);
        ly[j] = y0 + (y1 - y0) * rng.uniform(2, j);
    }
    // visibility through a uniform grid over the landmarks
    std::vector<int> nobs(NL);
    std::vector<int32_t> ebp, ebl;
    std::vector<double> ebz;
    for (int attempt = 0; attempt < 6; attempt++) {
        const double cell = range;
        const int gx = std::max(1, (int)std::ceil((x1 - x0) / cell)), gy = std::max(1, (int)std::ceil((y1 - y0) / cell));
        std::vector<int> head((size_t)gx * gy + 1, 0), items(NL);
        auto cell_of = [&](double x, double y) {
            int cx = std::min(gx - 1, std::max(0, (int)((x - x0) / cell)));
            int cy = std::min(gy - 1, std::max(0, (int)((y - y0) / cell)));
            return cy * gx + cx;
        };
        for (int j = 0; j < NL; j++) head[cell_of(lx[j], ly[j]) + 1]++;
        for (size_t c = 0; c < (size_t)gx * gy; c++) head[c + 1] += head[c];
        {
            std::vector<int> cur(head.begin(), head.end() - 1);
            for (int j = 0; j < NL; j++) items[cur[cell_of(lx[j], ly[j])]++] = j;
        }
        ebp.clear(); ebl.clear(); ebz.clear();
        std::fill(nobs.begin(), nobs.end(), 0);
        std::vector<std::pair<int, double>> seen;
        for (int i = 0; i < NP; i++) {
            const double c = std::cos(pth[i]), s = std::sin(pth[i]);
            const int cx = std::min(gx - 1, std::max(0, (int)((px[i] - x0) / cell)));
            const int cy = std::min(gy - 1, std::max(0, (int)((py[i] - y0) / cell)));
            seen.clear();
            for (int yy = std::max(0, cy - 1); yy <= std::min(gy - 1, cy + 1); yy++)
                for (int xx = std::max(0, cx - 1); xx <= std::min(gx - 1, cx + 1); xx++)
                    for (int q = head[yy * gx + xx]; q < head[yy * gx + xx + 1]; q++) {
                        const int j = items[q];
                        const double dx = lx[j] - px[i], dy = ly[j] - py[i];
                        const double d2 = dx * dx + dy * dy;
                        if (d2 > range * range || d2 < 0.04) continue;
                        const double gxr = c * dx + s * dy, gyr = -s * dx + c * dy;
                        const double bearing = std::atan2(gyr, gxr);
                        if (std::fabs(bearing) > 1.5707963267948966) continue;
                        seen.emplace_back(j, bearing);
                    }
            std::sort(seen.begin(), seen.end());  // ascending landmark inside a pose, like the bundled files
            for (auto& sb : seen) {
                ebp.push_back(i); ebl.push_back(sb.first); ebz.push_back(sb.second);
                nobs[sb.first]++;
            }
        }
        int deficient = 0;
        for (int j = 0; j < NL; j++)
            if (nobs[j] < 2) {
                deficient++;
                // re-seat the landmark ahead of a pose in the middle of a row: that pose and its row neighbours see it
                const int r = (int)(rng.uniform(10 + attempt, j) * rows) % rows;
                int k = 2 + (int)(rng.uniform(20 + attempt, j) * std::max(1, cols - 4));
                int i = std::min(NP - 1, r * cols + std::min(k, cols - 1));
                const double fwd = 0.6 * range * (0.3 + 0.6 * rng.uniform(30 + attempt, j));
                const double side = 0.3 * range * (rng.uniform(40 + attempt, j) - 0.5);
                const double th = (i / cols) % 2 == 0 ? 0.0 : 3.141592653589793;
                lx[j] = px[i] + std::cos(th) * fwd - std::sin(th) * side;
                ly[j] = py[i] + std::sin(th) * fwd + std::cos(th) * side;
            }
        if (deficient == 0) break;
    }
    // ---- ids ----------------------------------------------------------------------------------------------
    W->pose_ids.resize(NP);
    for (int i = 0; i < NP; i++) W->pose_ids[i] = 1200 + i;
    W->lm_ids.resize(NL);
    for (int j = 0; j < NL; j++) W->lm_ids[j] = 3 * j + (int)(splitmix64(spec->seed + j) % 3);  // ascending, with gaps
    // ---- measurements -------------------------------------------------------------------------------------
    const size_t Eb = ebp.size();
    W->b_pose_id.resize(Eb); W->b_lm_id.resize(Eb); W->b_z.resize(Eb);
    for (size_t e = 0; e < Eb; e++) {
        W->b_pose_id[e] = W->pose_ids[ebp[e]];
        W->b_lm_id[e] = W->lm_ids[ebl[e]];
        W->b_z[e] = f32(wrap(ebz[e] + spec->bearing_sigma * rng.normal(3, e)));
    }
    const int Eo = NP - 1;
    W->o_src_id.resize(Eo); W->o_dst_id.resize(Eo); W->o_z.resize(3 * (size_t)Eo); W->o_omega.assign(9 * (size_t)Eo, 0.0);
    const double wxy = f32(1.0 / (spec->odom_sigma_xy * spec->odom_sigma_xy));
    const double wth = f32(1.0 / (spec->odom_sigma_theta * spec->odom_sigma_theta));
    for (int e = 0; e < Eo; e++) {
        const int s = e, d = e + 1;
        W->o_src_id[e] = W->pose_ids[s]; W->o_dst_id[e] = W->pose_ids[d];
        const double c = std::cos(pth[s]), sn = std::sin(pth[s]);
        const double dx = px[d] - px[s], dy = py[d] - py[s];
        W->o_z[3 * (size_t)e] = f32(c * dx + sn * dy + spec->odom_sigma_xy * rng.normal(4, e));
        W->o_z[3 * (size_t)e + 1] = f32(-sn * dx + c * dy + spec->odom_sigma_xy * rng.normal(5, e));
        W->o_z[3 * (size_t)e + 2] = f32(wrap(pth[d] - pth[s] + spec->odom_sigma_theta * rng.normal(6, e)));
        W->o_omega[9 * (size_t)e] = wxy; W->o_omega[9 * (size_t)e + 4] = wxy; W->o_omega[9 * (size_t)e + 8] = wth;
    }
    // ---- states --------------------------------------------------------------------------------------------
    W->poses_true.resize(3 * (size_t)NP); W->poses_init.resize(3 * (size_t)NP);
    const double A = spec->init_drift, wn = spec->init_noise;
    const double P1 = NP / 3.0 + 1.0, P2 = NP / 2.3 + 1.0, P3 = NP / 4.1 + 1.0;
    for (int i = 0; i < NP; i++) {
        W->poses_true[3 * (size_t)i] = f32(px[i]); W->poses_true[3 * (size_t)i + 1] = f32(py[i]); W->poses_true[3 * (size_t)i + 2] = f32(pth[i]);
        double ddx = 0, ddy = 0, ddt = 0;
        if (i > 0) {
            ddx = A * std::sin(6.283185307179586 * i / P1) + wn * rng.normal(7, i);
            ddy = A * (1.0 - std::cos(6.283185307179586 * i / P2)) + wn * rng.normal(8, i);
            ddt = 0.1 * A * std::sin(6.283185307179586 * i / P3) + 0.1 * wn * rng.normal(9, i);
        }
        W->poses_init[3 * (size_t)i] = f32(px[i] + ddx);
        W->poses_init[3 * (size_t)i + 1] = f32(py[i] + ddy);
        W->poses_init[3 * (size_t)i + 2] = f32(wrap(pth[i] + ddt));
    }
    W->lms_true.resize(2 * (size_t)NL);
    for (int j = 0; j < NL; j++) { W->lms_true[2 * (size_t)j] = f32(lx[j]); W->lms_true[2 * (size_t)j + 1] = f32(ly[j]); }
    *out = W;
    return 0;
}

int bos_synth_destroy(bos_synth* w) {
    delete w;
    return 0;
}

int bos_synth_counts(const bos_synth* w, int64_t* c) {
    if (!w || !c) return 1;
    c[0] = (int64_t)w->pose_ids.size(); c[1] = (int64_t)w->lm_ids.size();
    c[2] = (int64_t)w->b_z.size(); c[3] = (int64_t)w->o_src_id.size();
    return 0;
}

int bos_synth_get(const bos_synth* w, int32_t* pose_ids, double* poses_xyt_init, double* poses_xyt_true, int32_t* lm_ids,
                  double* lms_xy_true, int32_t* b_pose_id, int32_t* b_lm_id, double* b_z, int32_t* o_src_id, int32_t* o_dst_id,
                  double* o_z, double* o_omega) {
    if (!w) return 1;
    auto cp = [](auto& v, auto* dst) { if (dst) std::copy(v.begin(), v.end(), dst); };
    cp(w->pose_ids, pose_ids); cp(w->poses_init, poses_xyt_init); cp(w->poses_true, poses_xyt_true);
    cp(w->lm_ids, lm_ids); cp(w->lms_true, lms_xy_true);
    cp(w->b_pose_id, b_pose_id); cp(w->b_lm_id, b_lm_id); cp(w->b_z, b_z);
    cp(w->o_src_id, o_src_id); cp(w->o_dst_id, o_dst_id); cp(w->o_z, o_z); cp(w->o_omega, o_omega);
    return 0;
}

}  // extern "C"
