/* Hand-written g2o build configuration for the parity oracle (test infrastructure).
 * Stands in for the file CMake would generate from thirdparty/g2o/config.h.in:
 * double precision, CSparse available, no OpenMP (the reference's shipped setting,
 * thirdparty/g2o/CMakeLists.txt:144), implicit ownership on. */
#ifndef G2O_CONFIG_H
#define G2O_CONFIG_H

#define G2O_HAVE_CSPARSE 1
#define G2O_DELETE_IMPLICITLY_OWNED_OBJECTS 1
#define G2O_NUMBER_FORMAT_STR "%lg"

#ifdef __cplusplus
using number_t = double;
#include <g2o/core/eigen_types.h>
#else
typedef double number_t;
#endif

#endif
