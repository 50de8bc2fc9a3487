// The six Schaefer-Turek configurations and their inlet profile, with the parameter values
// of reference src/classes/TestCases.hpp:14-89 (profile) and :97-308 (factories).
#pragma once
#include <cmath>
#include <memory>
#include <string>

#include "navier_stokes.hpp"

namespace nsb_host {

// 2-D: 4 U_m y (H-y)/H^2 in x;  3-D: 16 U_m x y (H-x)(H-y)/H^4 in z;  optional sin(pi t/8)
// modulation and half-cosine start-up ramp over T_ramp seconds.
template <int dim> class BenchmarkInletVelocity : public Function<dim> {
public:
  BenchmarkInletVelocity(double H_, double U_m_, bool time_dep, double T_ramp_ = 0.0)
    : Function<dim>(dim + 1), H(H_), U_m(U_m_), time_dependent(time_dep), T_ramp(T_ramp_) {}

  double value(const Point<dim>& p, const unsigned int component) const override {
    constexpr unsigned int flow_component = (dim == 2) ? 0 : 2;
    if (component != flow_component) return 0.0;
    double profile = 0.0;
    if (dim == 2) {
      const double y = p[1];
      profile = 4.0 * U_m * y * (H - y) / (H * H);
    } else {
      const double x = p[0], y = p[1];
      profile = 16.0 * U_m * x * y * (H - x) * (H - y) / (H * H * H * H);
    }
    const double t = this->get_time();
    if (time_dependent) profile *= std::sin(M_PI * t / 8.0);
    if (T_ramp > 0.0 && t < T_ramp) profile *= 0.5 * (1.0 - std::cos(M_PI * t / T_ramp));
    return profile;
  }

protected:
  const double H, U_m;
  const bool time_dependent;
  const double T_ramp;
};

namespace TestCases {

template <int dim>
inline BenchmarkTestCase<dim> make_case(const std::string& name, const std::string& description,
                                        const std::string& mesh_file, double Re, double U_m, double T, double deltat,
                                        TimeScheme ts, NonlinearMethod nm, bool time_dep, double t_ramp, bool supg) {
  constexpr double H = 0.41;
  BenchmarkTestCase<dim> tc;
  tc.name = name;
  tc.description = description;
  tc.mesh_file = mesh_file;
  tc.degree_velocity = 2;
  tc.degree_pressure = 1;
  tc.Re = Re;
  tc.U_m = U_m;
  tc.T = T;
  tc.deltat = deltat;
  tc.time_scheme = ts;
  tc.nonlinear_method = nm;
  tc.inlet_velocity = std::make_shared<BenchmarkInletVelocity<dim>>(H, U_m, time_dep, t_ramp);
  tc.dirichlet_bc = std::make_shared<ZeroDirichletBC<dim>>();
  tc.forcing_term = std::make_shared<ForcingTerm<dim>>();
  tc.initial_condition = std::make_shared<InitialCondition<dim>>();
  tc.use_supg = supg;
  return tc;
}

// 2D-1: steady, Re 20, U_m 0.3, BE + Newton, T 10, ramp 1 s          (TestCases.hpp:101-131)
inline BenchmarkTestCase<2> make_2D_1(const std::string& mesh_file, TimeScheme ts = TimeScheme::BackwardEuler,
                                      NonlinearMethod nm = NonlinearMethod::Newton, double deltat = -1.0,
                                      double t_ramp = 1.0) {
  return make_case<2>("2D-1", "Steady flow around cylinder, Re=20, U_m=0.3", mesh_file, 20.0, 0.3, 10.0, deltat, ts, nm,
                      false, t_ramp, false);
}
// 2D-2: unsteady, Re 100, U_m 1.5, CN + linearised, T 8, ramp 2 s    (TestCases.hpp:134-168)
inline BenchmarkTestCase<2> make_2D_2(const std::string& mesh_file, TimeScheme ts = TimeScheme::CrankNicolson,
                                      NonlinearMethod nm = NonlinearMethod::Linearized, double deltat = -1.0) {
  return make_case<2>("2D-2", "Unsteady flow, Re=100, U_m=1.5, constant inlet", mesh_file, 100.0, 1.5, 8.0, deltat, ts,
                      nm, false, 2.0, false);
}
// 2D-3: sin(pi t/8) inlet                                            (TestCases.hpp:171-201)
inline BenchmarkTestCase<2> make_2D_3(const std::string& mesh_file, TimeScheme ts = TimeScheme::CrankNicolson,
                                      NonlinearMethod nm = NonlinearMethod::Linearized, double deltat = -1.0) {
  return make_case<2>("2D-3", "Unsteady flow, time-varying inlet sin(pi*t/8), U_m=1.5, Re(t) in [0,100]", mesh_file,
                      100.0, 1.5, 8.0, deltat, ts, nm, true, 0.0, false);
}
// 3D-1Z: steady, Re 20, U_m 0.45, BE + Newton, SUPG                  (TestCases.hpp:204-234)
inline BenchmarkTestCase<3> make_3D_1Z(const std::string& mesh_file, TimeScheme ts = TimeScheme::BackwardEuler,
                                       NonlinearMethod nm = NonlinearMethod::Newton, double deltat = -1.0) {
  return make_case<3>("3D-1Z", "Steady 3D flow, Re=20, U_m=0.45, circular cylinder", mesh_file, 20.0, 0.45, 10.0,
                      deltat, ts, nm, false, 0.0, true);
}
// 3D-2Z: unsteady, Re 100, U_m 2.25, CN + linearised, dt 0.01, ramp 4 s, SUPG (TestCases.hpp:237-270)
inline BenchmarkTestCase<3> make_3D_2Z(const std::string& mesh_file, TimeScheme ts = TimeScheme::CrankNicolson,
                                       NonlinearMethod nm = NonlinearMethod::Linearized, double deltat = -1.0) {
  return make_case<3>("3D-2Z", "Unsteady 3D flow, Re=100, U_m=2.25, circular cylinder, constant inlet", mesh_file,
                      100.0, 2.25, 8.0, (deltat > 0) ? deltat : 0.01, ts, nm, false, 4.0, true);
}
// 3D-3Z: sin(pi t/8) inlet, dt 0.01, SUPG                            (TestCases.hpp:273-306)
inline BenchmarkTestCase<3> make_3D_3Z(const std::string& mesh_file, TimeScheme ts = TimeScheme::CrankNicolson,
                                       NonlinearMethod nm = NonlinearMethod::Linearized, double deltat = -1.0) {
  return make_case<3>("3D-3Z", "Unsteady 3D flow, time-varying inlet sin(pi*t/8), U_m=2.25, Re(t) in [0,100]",
                      mesh_file, 100.0, 2.25, 8.0, (deltat > 0) ? deltat : 0.01, ts, nm, true, 0.0, true);
}

}  // namespace TestCases
}  // namespace nsb_host
