/* Shim for the public header the vendored secp256k1 sources expect at "../include/secp256k1.h"
 * (/root/reference/porla/Utils/secp256k1_lib/secp256k1.c:9, precomputed_ecmult.c:8).  The real
 * header is installed system-wide in a Porla build (-I/usr/local/include, porla/Makefile:3) and is
 * absent here; only these macros and tag constants are needed by the *_impl.h files on the
 * ecmult_multi path.  TEST INFRASTRUCTURE ONLY. */
#ifndef SECP256K1_H
#define SECP256K1_H
#include <stddef.h>
#if !defined(SECP256K1_GNUC_PREREQ)
# if defined(__GNUC__) && defined(__GNUC_MINOR__)
#  define SECP256K1_GNUC_PREREQ(_maj,_min) ((__GNUC__<<16)+__GNUC_MINOR__>=((_maj)<<16)+(_min))
# else
#  define SECP256K1_GNUC_PREREQ(_maj,_min) 0
# endif
#endif
#define SECP256K1_INLINE inline
#define SECP256K1_API
#define SECP256K1_WARN_UNUSED_RESULT __attribute__((__warn_unused_result__))
#define SECP256K1_ARG_NONNULL(_x)
#define SECP256K1_TAG_PUBKEY_EVEN 0x02
#define SECP256K1_TAG_PUBKEY_ODD 0x03
#define SECP256K1_TAG_PUBKEY_UNCOMPRESSED 0x04
#define SECP256K1_TAG_PUBKEY_HYBRID_EVEN 0x06
#define SECP256K1_TAG_PUBKEY_HYBRID_ODD 0x07
#endif
