/* Drop-in for "ectrans/version.h" (src/transi/version.h); the declarations live in transi_b200.h. */
#include "../transi_b200.h"
