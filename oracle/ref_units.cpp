// ref_units.cpp — TEST INFRASTRUCTURE (oracle/): the REFERENCE's unit system and constants run here.
// shamunits/include/shamunits/{UnitSystem,Constants}.hpp are plain C++; this driver is compiled against them where they
// lie (oracle/Makefile, target _ref/units_ref).  The disc configuration (examples/sph/run_circular_disc_central_pot.py)
// takes its gravitational constant from them.
//   units_ref unit_time unit_length unit_mass   ->  G year au sol_mass   (in those code units, %.17g)
#include "shamunits/Constants.hpp"
#include "shamunits/UnitSystem.hpp"
#include <cstdio>
#include <cstdlib>

int main(int argc, char **argv) {
    if (argc != 4)
        return 1;
    shamunits::UnitSystem<double> u(std::atof(argv[1]), std::atof(argv[2]), std::atof(argv[3]));
    shamunits::Constants<double> c(u);
    std::printf("%.17g %.17g %.17g %.17g\n", c.G(), c.year(), c.au(), c.sol_mass());
    return 0;
}
