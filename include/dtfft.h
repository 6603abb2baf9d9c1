/* Drop-in name for sources that `#include <dtfft.h>`: the API itself is declared in
 * dtfft_b200_api.h (see its header comment for the two differences from the reference). */
#ifndef DTFFT_H
#define DTFFT_H
#include "dtfft_b200_api.h"
#define DTFFT_CALL(call)                                                                              \
    do {                                                                                              \
        dtfft_error_t ierr_ = (call);                                                                 \
        if (ierr_ != DTFFT_SUCCESS) {                                                                 \
            fprintf(stderr, "dtFFT error in file '%s:%i': %s.\n", __FILE__, __LINE__,                \
                    dtfft_get_error_string(ierr_));                                                   \
            abort();                                                                                  \
        }                                                                                             \
    } while (0)
#include <stdio.h>
#include <stdlib.h>
#endif
