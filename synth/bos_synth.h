/* bos_synth.h -- synthetic bearing-only worlds: the INPUT GENERATOR of the benchmark and of the tests (host code, C ABI).
 *
 * Not part of the product library: it lives in its own shared object (synth/libbos_synth.so) so that the CPU reference arm of
 * bench.py generates its workload without mapping libbos_b200.so.  Mirrors the statistics of the reference's bundled dataset
 * (data/slam2D_bearing_only_*.g2o), see bos_synth.cpp.
 */
#ifndef BOS_SYNTH_H
#define BOS_SYNTH_H

#include <stdint.h>

#if defined(__GNUC__)
#define BOS_SYNTH_API __attribute__((visibility("default")))
#else
#define BOS_SYNTH_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bos_synth_spec {
    int n_poses;
    int n_landmarks;
    int64_t target_bearing_edges;   /* sensor range is tuned to approach this count */
    uint64_t seed;
    double bearing_sigma;           /* default 3e-3 rad */
    double odom_sigma_xy;           /* default 1/sqrt(500) */
    double odom_sigma_theta;        /* default 1/sqrt(5000) */
    double init_drift;              /* amplitude (m) of the smooth drift applied to the ground truth for the initial guess */
    double init_noise;              /* white noise (m, rad/10) on the initial guess */
    int reserved[8];
} bos_synth_spec;
typedef struct bos_synth bos_synth;
BOS_SYNTH_API void bos_synth_default_spec(bos_synth_spec* s);
BOS_SYNTH_API int bos_synth_create(const bos_synth_spec* spec, bos_synth** out);
BOS_SYNTH_API int bos_synth_destroy(bos_synth* w);
/* counts[0..3] = NP, NL, Eb, Eo */
BOS_SYNTH_API int bos_synth_counts(const bos_synth* w, int64_t* counts4);
/* ids and values of the generated world; every pointer may be NULL.  Values are rounded to float and
 * widened, as the g2o loader does (utils/g2o_utils.cpp: std::stof). */
BOS_SYNTH_API int bos_synth_get(const bos_synth* w, int32_t* pose_ids, double* poses_xyt_init, double* poses_xyt_true,
                  int32_t* lm_ids, double* lms_xy_true,
                  int32_t* b_pose_id, int32_t* b_lm_id, double* b_z,
                  int32_t* o_src_id, int32_t* o_dst_id, double* o_z, double* o_omega);

#ifdef __cplusplus
}
#endif
#endif /* BOS_SYNTH_H */
