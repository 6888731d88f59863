/* Lets strom's `#include "libhmsbeagle/beagle.h"` (reference src/likelihood.hpp:8, src/model.hpp:5)
 * resolve to the B200 engine's boundary header: add -I<repo>/include/compat and link -lstrom_b200. */
#include "../../strom_beagle.h"
