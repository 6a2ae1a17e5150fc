// Closest isotropic tensor in the Frobenius norm (mirrors TensorProjection.hh:20-53): project C onto the
// span of the hydrostatic projector J and the deviatoric projector K,
//   alpha = C:J / J:J,  beta = C:K / K:K,  lambda = (alpha - beta)/N,  mu = beta/2.
#ifndef MESHFEM_B200_TENSORPROJECTION_HH
#define MESHFEM_B200_TENSORPROJECTION_HH
#include <MeshFEM/ElasticityTensor.hh>

template <typename Real, size_t N>
ElasticityTensor<Real, N> closestIsotropicTensor(const ElasticityTensor<Real, N> &C) {
    Real C_ijij = 0.0, C_iijj = 0.0;
    for (size_t i = 0; i < N; ++i)
        for (size_t j = 0; j < N; ++j) {
            C_ijij += C(i, j, i, j);
            C_iijj += C(i, i, j, j);
        }
    const Real n = N;
    const Real CdotJ = C_iijj / n, CdotK = C_ijij - CdotJ;
    const Real KdotK = 0.5 * (n * n + n) - 1.0;
    const Real alpha = CdotJ, beta = CdotK / KdotK;
    ElasticityTensor<Real, N> result;
    result.setIsotropicLame((alpha - beta) / n, beta / 2.0);
    return result;
}
#endif
