/* Stand-in for <ceres/ceres.h> (Ceres is not installed): the base classes LidarFactor and ResidualBlockInfo use
 * (lidar_factor.h:31, residual_block_info.h) and HuberLoss written from Ceres' documented definition
 * (loss_function.h: rho(s) = s for s <= delta^2, 2 delta sqrt(s) - delta^2 beyond). The corrector that USES rho is the
 * reference's own code (residual_block_info.cc:49-77); the three numbers HuberLoss returns are modelled here. */
#ifndef FFLINS_ORACLE_REFSHIM_CERES_H
#define FFLINS_ORACLE_REFSHIM_CERES_H
#include "../mini_eigen.h"
#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>
namespace ceres {
class CostFunction {
public:
    virtual ~CostFunction() {}
    virtual bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const = 0;
    const std::vector<int> &parameter_block_sizes() const { return sizes_; }
    int num_residuals() const { return num_residuals_; }
protected:
    std::vector<int> *mutable_parameter_block_sizes() { return &sizes_; }
    void set_num_residuals(int n) { num_residuals_ = n; }
    std::vector<int> sizes_;
    int num_residuals_ = 0;
};
template <int kNumResiduals, int... Ns> class SizedCostFunction : public CostFunction {
public:
    SizedCostFunction() {
        num_residuals_ = kNumResiduals;
        sizes_         = std::vector<int>{Ns...};
    }
};
class Manifold {
public:
    virtual ~Manifold() {}
    virtual int AmbientSize() const = 0;
    virtual int TangentSize() const = 0;
    virtual bool Plus(const double *x, const double *delta, double *x_plus_delta) const = 0;
    virtual bool PlusJacobian(const double *x, double *jacobian) const = 0;
    virtual bool Minus(const double *y, const double *x, double *y_minus_x) const = 0;
    virtual bool MinusJacobian(const double *x, double *jacobian) const = 0;
};
class LossFunction {
public:
    virtual ~LossFunction() {}
    virtual void Evaluate(double sq_norm, double out[3]) const = 0;
};
class HuberLoss : public LossFunction {
    const double a_, b_;
public:
    explicit HuberLoss(double a) : a_(a), b_(a * a) {}
    void Evaluate(double s, double rho[3]) const override {
        if (s > b_) {
            const double r = std::sqrt(s);
            rho[0]         = 2.0 * a_ * r - b_;
            rho[1]         = std::max(std::numeric_limits<double>::min(), a_ / r);
            rho[2]         = -rho[1] / (2.0 * s);
        } else {
            rho[0] = s;
            rho[1] = 1.0;
            rho[2] = 0.0;
        }
    }
};
} // namespace ceres
#endif
