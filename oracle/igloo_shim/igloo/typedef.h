/* oracle/igloo_shim/igloo/typedef.h -- TEST INFRASTRUCTURE (see ro.h).
 * The reference's include/coolmic-dsp/types.h:33-50 forward-declares its object
 * types through these two macros; with an untyped igloo_ro_t they are no-ops. */
#ifndef ORACLE_IGLOO_SHIM_TYPEDEF_H
#define ORACLE_IGLOO_SHIM_TYPEDEF_H
#define igloo_RO_FORWARD_TYPE(T)  struct igloo_shim_forward_decl_unused
#define igloo_RO_TYPE(T)
#endif
