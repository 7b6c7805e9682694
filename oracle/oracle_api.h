/*
 * oracle_api.h — C ABI shared by the two CPU checkers.  TEST INFRASTRUCTURE ONLY.
 *
 *   oracle/libmcl3dl_oracle.so      "port"      : repo-owned restatement (oracle/mcl3dl_oracle.cpp)
 *   oracle/_ref/libmcl3dl_ref.so    "reference" : the reference's own unmodified sources compiled
 *                                                 against oracle/shim/ (oracle/ref_driver.cpp)
 *
 * Both export exactly these symbols so tests can run the same vectors through either.  Nothing
 * in the product path (mcl_3dl_b200/) may include, link or load this.
 */
#ifndef MCL3DL_ORACLE_API_H
#define MCL3DL_ORACLE_API_H

#include "../include/mcl3dl_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* LidarMeasurementModelBeamParameters as the node fills it (include/mcl_3dl/parameters.h:91-132). */
typedef struct
{
  float map_grid_x, map_grid_y, map_grid_z;
  uint64_t num_points_default;
  float beam_likelihood_min;
  float ang_total_ref;
  uint32_t filter_label_max;
  float hit_range;
  int32_t add_penalty_short_only_mode;
  int32_t use_raycast_using_dda; /* must be 1 for parity with the GPU beam kernel */
  float ray_angle_half;
  float dda_grid_size;
} mcl3dl_cpu_beam_raw;

typedef struct mcl3dl_cpu mcl3dl_cpu;

const char* mcl3dl_cpu_kind(void); /* "port" or "reference" */

/* Build the CPU map index (ChunkedKdtree(chunk_length, max_search_radius) + setInputCloud, and a
 * LidarMeasurementModelBeam / Likelihood pair).  lik or beam may be NULL -> reference defaults. */
mcl3dl_cpu* mcl3dl_cpu_create(const mcl3dl_point* pts, size_t n, const mcl3dl_lik_params* lik,
                              const mcl3dl_cpu_beam_raw* beam, float chunk_length, float max_search_radius);
void mcl3dl_cpu_destroy(mcl3dl_cpu*);

/* pf.measure()'s per-particle body (mcl_3dl.cpp:409-415) for P particles; n_threads > 1 is the
 * "not reference behaviour" all-core variant (one raycaster per thread). */
int mcl3dl_cpu_measure(mcl3dl_cpu*, const mcl3dl_pose* poses, size_t n_particles,
                       const mcl3dl_point* lik_pts, size_t n_lik,
                       const mcl3dl_point* beam_pts, size_t n_beam,
                       const float* origins_xyz, size_t n_origins, mcl3dl_result* out, int n_threads);

/* enable (default) / disable the extra per-ray getBeamStatus pass that fills n_short/n_hit/n_long;
 * bench.py disables it so the timed CPU work is exactly pf.measure()'s. */
int mcl3dl_cpu_set_tally(mcl3dl_cpu*, int enable);

/* BeamStatus per (particle, ray), row-major [P][n_beam]: 0 SHORT, 1 HIT, 2 LONG, 3 TOTAL_REFLECTION. */
int mcl3dl_cpu_beam_status(mcl3dl_cpu*, const mcl3dl_pose* poses, size_t n_particles,
                           const mcl3dl_point* beam_pts, size_t n_beam,
                           const float* origins_xyz, size_t n_origins, uint8_t* status);

/* The derived beam parameters (refreshParameters) this checker uses. */
int mcl3dl_cpu_beam_params(mcl3dl_cpu*, mcl3dl_beam_params* out);

/* ChunkedKdtree::radiusSearch(q, radius, id, d2, 1) -> original id or -1; *d2 in the rescaled space. */
int mcl3dl_cpu_radius_search(mcl3dl_cpu*, const float q[3], float radius, float* d2);

/* Stand-alone RaycastUsingDDA<PointXYZ>(gx,gy,gz,dda,ray_angle_half,hit_tol) over
 * ChunkedKdtree(10.0, 1.0): setRay(begin,end) then getNextCastResult until exhausted
 * (or until the first collision when stop_at_collision).  Writes voxel centres (3 floats each) and
 * collision flags; returns the number of cast results (<= max_out) or <0 on error.
 * collided_id (may be NULL) receives the map index of the first colliding point or -1. */
int mcl3dl_cpu_dda_walk(const mcl3dl_point* pts, size_t n, const double ctor[6],
                        const float begin[3], const float end[3], int stop_at_collision,
                        float* centres, uint8_t* collision, int max_out, int* collided_id);

/* Stand-alone RaycastUsingKDTree<PointXYZ>(gx,gy,gz,hit_tol) over ChunkedKdtree(10.0, 1.0)
 * (test/src/test_raycast.cpp): positions (pos_), collision flags and sin_angle_ of every cast result. */
int mcl3dl_cpu_kd_walk(const mcl3dl_point* pts, size_t n, const float ctor[4], const float begin[3], const float end[3],
                       int stop_at_collision, float* positions, uint8_t* collision, float* sin_angle, int max_out,
                       int* collided_id);

/* Quat * Vec3 (quat.h:139-143) and State6DOF::transform of one point (state_6dof.h:214-225). */
void mcl3dl_cpu_quat_rotate(const float q[4], const float v[3], float out[3]);
void mcl3dl_cpu_transform_point(const mcl3dl_pose* pose, const float v[3], float out[3]);

/* pf::ParticleFilter::measure's weight update (pf.h:252-279) given per-particle likelihood values:
 * prob[i] *= lik[i]; normalise; entropy; returns 1 if sum > 0 else 0 (restore -> prob untouched). */
int mcl3dl_cpu_pf_update(float* prob, const float* lik, size_t n, float* entropy);

/* Groundwork for scope row f3 (not on the device yet): MotionPredictionModelDifferentialDrive::setOdoms + predict
 * (include/mcl_3dl/motion_prediction_models/motion_prediction_model_differential_drive.h:46-67) applied to n states. */
typedef struct
{
  float pos[3];
  float rot[4]; /* x y z w */
  float noise_ll, noise_la, noise_al, noise_aa;
  float odom_err_integ_lin[3];
  float odom_err_integ_ang[3];
} mcl3dl_cpu_motion_state; /* 17 floats: the State6DOF fields predict() reads and writes (state_6dof.h:55-63) */
int mcl3dl_cpu_motion_predict(const mcl3dl_pose* odom_prev, const mcl3dl_pose* odom_current, float time_diff,
                              float odom_err_integ_lin_tc, float odom_err_integ_ang_tc,
                              mcl3dl_cpu_motion_state* states, size_t n);

/* Groundwork for scope row f4: the models' filter() without the sampler (identical code in both models,
 * src/lidar_measurement_model_likelihood.cpp:79-103 / _beam.cpp:98-122): keep[i] = 1 if point i survives the clip by
 * planar range and z window.  num_points_after = the sample size setGlobalLocalizationStatus (:63-77) would request
 * for (num_points_default, num_points_global, num_particles, current_num_particles). */
int mcl3dl_cpu_filter_clip(const mcl3dl_point* pts, size_t n, float clip_near, float clip_far, float clip_z_min,
                           float clip_z_max, uint8_t* keep);
size_t mcl3dl_cpu_global_localization_points(size_t num_points_default, size_t num_points_global, size_t num_particles,
                                             size_t current_num_particles);

/* pf::ParticleFilter<State1D, float>(n, seed)::resample(State1D(sigma)) (pf.h:182-225) on 1-D states, the fixture
 * of test/src/test_pf.cpp:186-289: systematic resampling over the sorted cumulative weights, noise only on
 * duplicates, std::default_random_engine.  Writes the resampled states and probabilities. */
int mcl3dl_cpu_pf_resample_1d(const float* probs, const float* states, size_t n, unsigned int seed, float sigma,
                              float* out_states, float* out_probs);

/* pf::ParticleFilter<State6DOF, float, ParticleWeightedMeanQuat, std::default_random_engine>(n, seed)
 *   ::resample(State6DOF(sigma_pos, sigma_rpy))   (pf.h:182-225 with State6DOF::generateNoise / operator+,
 * state_6dof.h:226-261): what the node runs after every measurement (src/mcl_3dl.cpp:809-815).  In/out: the
 * 17-float states above (duplicates come back with noise_* = 0, as `ret` is a fresh State6DOF) and probabilities. */
int mcl3dl_cpu_pf_resample_6dof(const float* probs, const mcl3dl_cpu_motion_state* states, size_t n, unsigned int seed,
                                const float sigma_pos[3], const float sigma_rpy[3], mcl3dl_cpu_motion_state* out_states,
                                float* out_probs);

/* The pose estimate the node takes from the filter after every measurement (src/mcl_3dl.cpp:428-452,704-724):
 *   pf_->bias(f)  with f = NormalLikelihood(bias_var_dist)(|pos - prev.pos|) * NormalLikelihood(bias_var_ang)(angle of
 *                 rot * prev.rot.inv()) + 1e-6   (state_prev == NULL: the constant 1 of the global-localisation branch),
 *   pf_->expectationBiased()  (pf.h:294-303, ParticleWeightedMeanQuat state_6dof.h:316-355),
 *   pf_->max()                (pf.h:361-374: first particle with the largest probability_),
 *   pf_->covariance(1.0, 1.0) (pf.h:304-360 around expectation(1.0); State6DOF::covElement state_6dof.h:162-184).
 * mean_biased: pos + rot as returned (NOT normalised; the node normalises afterwards, mcl_3dl.cpp:462). */
int mcl3dl_cpu_pf_estimate(const float* probs, const mcl3dl_cpu_motion_state* states, size_t n,
                           const mcl3dl_pose* state_prev, float bias_var_dist, float bias_var_ang,
                           mcl3dl_pose* mean_biased, uint32_t* max_index, float cov[36]);

#ifdef __cplusplus
}
#endif
#endif
