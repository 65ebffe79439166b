/* GSL-subset compat: see gsl_compat_core.h */
#include "gsl_compat_core.h"
