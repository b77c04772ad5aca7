/* h1v2_oracle.h -- C interface of the CPU oracle (test infrastructure; see h1v2_oracle.c header).
 * Mirrors include/h1v2_b200.h one-to-one with HOST pointers so parity tests read alike. */
#ifndef H1V2_ORACLE_H
#define H1V2_ORACLE_H
#include <stdint.h>

#include "../include/h1v2_b200.h"

#ifdef __cplusplus
extern "C" {
#endif
typedef struct H1v2Oracle H1v2Oracle;

int h1v2o_create(const H1v2Config* cfg, int32_t n_envs, uint64_t seed, H1v2Oracle** out);
void h1v2o_destroy(H1v2Oracle* o);
void h1v2o_set_threads(H1v2Oracle* o, int n);
int h1v2o_obs_dim(const H1v2Oracle* o);
int64_t h1v2o_max_episode_length(const H1v2Oracle* o);
int h1v2o_reset(H1v2Oracle* o, const int64_t* env_ids, int32_t n);
int h1v2o_observe(H1v2Oracle* o, float* obs);
int h1v2o_step(H1v2Oracle* o, const float* actions, float* obs, float* rew, uint8_t* terminated, uint8_t* truncated);
int h1v2o_step_injected(H1v2Oracle* o, const float* actions, const float* qpos, const float* qvel, const float* timers,
                        const float* slot_hist, const float* tau, const float* qacc, const float* foot_vel, float* obs, float* rew,
                        uint8_t* terminated, uint8_t* truncated);
int h1v2o_get_state(H1v2Oracle* o, const H1v2State* s);
int h1v2o_set_state(H1v2Oracle* o, const H1v2State* s);
int h1v2o_get_episode_length(H1v2Oracle* o, int64_t* out);
int h1v2o_set_episode_length(H1v2Oracle* o, const int64_t* in);
int h1v2o_get_log(H1v2Oracle* o, float* out);
int h1v2o_set_reward_weights(H1v2Oracle* o, const float* w);
int h1v2o_solver_stats(H1v2Oracle* o, int32_t* iters, double* resid);
int h1v2o_activation_margin(H1v2Oracle* o, double* contact, double* limit);
/* rough terrain */
int h1v2o_terrain_dims(const H1v2Oracle* o, int32_t dims[2]);
int h1v2o_get_terrain(const H1v2Oracle* o, float* out);
int h1v2o_set_terrain(H1v2Oracle* o, const float* in);
float h1v2o_terrain_level_mean(const H1v2Oracle* o);
int h1v2o_terrain_query(const H1v2Oracle* o, int level, int type, double lx, double ly, double* h, double* n);
int h1v2o_tri_margin(H1v2Oracle* o, double* out);

/* building blocks for known-answer tests */
void h1v2o_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void h1v2o_rng4(uint64_t seed, int64_t env_gid, uint64_t step, uint32_t stream, uint32_t block, float u[4]);
void h1v2o_fk(const double* qpos, double* R_out, double* x_out);
void h1v2o_mass_matrix(const H1v2Config* cfg, const double* qpos, double* M);
void h1v2o_bias(const H1v2Config* cfg, const double* qpos, const double* qvel, double* bias);
void h1v2o_physics_step(const H1v2Config* cfg, double* qpos, double* qvel, const double* ctrl, double friction,
                        double* slot_force, int32_t* iters, double* resid);
double h1v2o_total_energy(const H1v2Config* cfg, const double* qpos, const double* qvel);
float h1v2o_wrap_to_pi(float a);
#ifdef __cplusplus
}
#endif
#endif
