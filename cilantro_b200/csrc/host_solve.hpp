// Host-side O(1) solves (product code). See host_solve.cpp.
#pragma once

namespace cb {

void t34_identity(float* T);
bool kabsch_from_moments(const double* sums16, float* T);
// the same from moments taken about the pivots pd (dst side) and pq (transformed-src side)
bool kabsch_from_pivoted_moments(const double* sums16, const float* pd, const float* pq, float* T);
bool gauss_newton_update(const double* sums28, const float* Tin, float* Tout, float* dtheta_norm);
void uncenter(float* T, const float* dst_mean, const float* src_mean);
void reorthonormalize(float* T);
void compose(const float* A, const float* B, float* out);
float update_norm(const float* T);
void apply_point(const float* T, const float* p, float* q);
void pca_from_cov(const double* cov9, float* evals3, float* evecs9);

}  // namespace cb
