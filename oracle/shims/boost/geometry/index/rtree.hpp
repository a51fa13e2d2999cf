// Shim: everything lives in boost/geometry.hpp (test infrastructure only, see oracle/README.md)
#include "../../geometry.hpp"
