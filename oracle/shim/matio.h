/* TEST INFRASTRUCTURE ONLY (oracle shim) - stands in for <matio.h>.  Mat_CreateVer returns NULL, so the
 * reference's .mat dumps are skipped (they test matfp before writing); binary dumps are unaffected. */
#pragma once
#include <cstddef>
typedef struct mat_t mat_t;
typedef struct matvar_t matvar_t;
enum mat_ft
{
    MAT_FT_MAT73 = 0x0200,
    MAT_FT_MAT5 = 0x0100
};
enum matio_classes
{
    MAT_C_DOUBLE = 6,
    MAT_C_SINGLE = 7,
    MAT_C_INT8 = 8,
    MAT_C_UINT8 = 9,
    MAT_C_INT16 = 10,
    MAT_C_UINT16 = 11,
    MAT_C_INT32 = 12,
    MAT_C_UINT32 = 13,
    MAT_C_INT64 = 14,
    MAT_C_UINT64 = 15
};
enum matio_types
{
    MAT_T_INT8 = 1,
    MAT_T_UINT8 = 2,
    MAT_T_INT16 = 3,
    MAT_T_UINT16 = 4,
    MAT_T_INT32 = 5,
    MAT_T_UINT32 = 6,
    MAT_T_SINGLE = 7,
    MAT_T_DOUBLE = 9,
    MAT_T_INT64 = 12,
    MAT_T_UINT64 = 13
};
enum matio_compression
{
    MAT_COMPRESSION_NONE = 0,
    MAT_COMPRESSION_ZLIB = 1
};
static inline mat_t* Mat_CreateVer(const char*, const char*, enum mat_ft) { return NULL; }
static inline int Mat_Close(mat_t*) { return 0; }
static inline matvar_t* Mat_VarCreate(const char*, enum matio_classes, enum matio_types, int, size_t*, const void*, int) { return NULL; }
static inline int Mat_VarWrite(mat_t*, matvar_t*, enum matio_compression) { return 0; }
static inline void Mat_VarFree(matvar_t*) {}
