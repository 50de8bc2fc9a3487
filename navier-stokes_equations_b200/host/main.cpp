// Driver with the same role as the reference's src/main.cpp:3-46, except that the test case and
// mesh are chosen on the command line instead of by editing the source:
//   navier_stokes <test case> <mesh file> [--steps N] [--no-vtu] [--tol X] [--device D]
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "navier_stokes.hpp"
#include "test_cases.hpp"

using namespace nsb_host;

int main(int argc, char* argv[]) {
  try {
    if (argc < 3) {
      std::cerr << "usage: " << argv[0] << " <2D-1|2D-2|2D-3|3D-1Z|3D-2Z|3D-3Z> <mesh.msh|mesh.bin> [--steps N] [--no-vtu] [--tol X] [--device D]" << std::endl;
      return 2;
    }
    const std::string tc_name = argv[1], mesh = argv[2];
    RunOptions opt;
    for (int i = 3; i < argc; ++i) {
      if (!std::strcmp(argv[i], "--steps") && i + 1 < argc) opt.max_steps = std::atoi(argv[++i]);
      else if (!std::strcmp(argv[i], "--no-vtu")) opt.write_vtu = false;
      else if (!std::strcmp(argv[i], "--tol") && i + 1 < argc) opt.gmres_tolerance = std::atof(argv[++i]);
      else if (!std::strcmp(argv[i], "--device") && i + 1 < argc) opt.device = std::atoi(argv[++i]);
    }
    if (tc_name == "2D-1") { NavierStokes<2> s(TestCases::make_2D_1(mesh), opt); s.run(); }
    else if (tc_name == "2D-2") { NavierStokes<2> s(TestCases::make_2D_2(mesh), opt); s.run(); }
    else if (tc_name == "2D-3") { NavierStokes<2> s(TestCases::make_2D_3(mesh), opt); s.run(); }
    else if (tc_name == "3D-1Z") { NavierStokes<3> s(TestCases::make_3D_1Z(mesh), opt); s.run(); }
    else if (tc_name == "3D-2Z") { NavierStokes<3> s(TestCases::make_3D_2Z(mesh), opt); s.run(); }
    else if (tc_name == "3D-3Z") { NavierStokes<3> s(TestCases::make_3D_3Z(mesh), opt); s.run(); }
    else { std::cerr << "unknown test case " << tc_name << std::endl; return 2; }
  } catch (std::exception& exc) {
    std::cerr << std::endl << "----------------------------------------------------" << std::endl;
    std::cerr << "Exception: " << exc.what() << std::endl << "Aborting!" << std::endl;
    return 1;
  } catch (...) {
    std::cerr << std::endl << "----------------------------------------------------" << std::endl;
    std::cerr << "Unknown exception! Aborting!" << std::endl;
    return 1;
  }
  return 0;
}
