/* oracle/igloo_shim/igloo/types.h -- TEST INFRASTRUCTURE (see ro.h).
 * Real libigloo builds a transparent union out of igloo_RO_APPTYPES here; the
 * shim's igloo_ro_t is void * so nothing is required. */
#ifndef ORACLE_IGLOO_SHIM_TYPES_H
#define ORACLE_IGLOO_SHIM_TYPES_H
#include "typedef.h"
#endif
