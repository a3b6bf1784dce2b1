/* include/compat/binmat.h -- drop-in for the reference header of the same name:
 * everything it declared is provided by liboblas_b200.so through oblas_compat.h. */
#include "../oblas_compat.h"
