// TEST INFRASTRUCTURE ONLY (oracle/): TF_DISALLOW_COPY_AND_ASSIGN as used at
// ctc_ext_beam_search_decoder.h:63 and ctc_beam_entry.h:245.
#ifndef CTCX_ORACLE_SHIM_MACROS_H_
#define CTCX_ORACLE_SHIM_MACROS_H_
#define TF_DISALLOW_COPY_AND_ASSIGN(TypeName) \
  TypeName(const TypeName&) = delete;         \
  void operator=(const TypeName&) = delete
#endif
