/* Stand-in for <ceres/ceres.h>: only the base class LidarFactor derives from (lidar_factor.h:31). */
#ifndef FFLINS_ORACLE_REFSHIM_CERES_H
#define FFLINS_ORACLE_REFSHIM_CERES_H
namespace ceres {
class CostFunction {
public:
    virtual ~CostFunction() {}
    virtual bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const = 0;
};
template <int kNumResiduals, int... Ns> class SizedCostFunction : public CostFunction {};
} // namespace ceres
#endif
