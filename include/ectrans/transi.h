/* Drop-in for the reference's "ectrans/transi.h" (src/transi/transi.h): a transi program compiles unchanged with
 * -I<repo>/include and links libectrans_b200.so. */
#include "../transi_b200.h"
