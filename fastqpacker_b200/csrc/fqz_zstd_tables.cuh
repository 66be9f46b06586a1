// fqz_zstd_tables.cuh — RFC 8878 constants shared by the GPU zstd encoder and decoder.
// (The reference's entropy stage is klauspost/compress/zstd v1.19.1, go.mod:8, not in the tree;
//  these tables are the format's published constants, validated by round trips through libzstd.)
#pragma once
#include "fqz_common.cuh"

#define ZSTD_MAGIC 0xFD2FB528u
#define ZSTD_BLOCK_MAX (128u * 1024u)
#define ZSTD_LL_MAXLOG 9
#define ZSTD_ML_MAXLOG 9
#define ZSTD_OF_MAXLOG 8
#define ZSTD_LL_DEFLOG 6
#define ZSTD_ML_DEFLOG 6
#define ZSTD_OF_DEFLOG 5
#define ZSTD_MAX_LL 35
#define ZSTD_MAX_ML 52
#define ZSTD_MAX_OF 31
#define HUF_MAXBITS 11

// extra bits and base values of literal-length / match-length codes
static __constant__ u8 kLLBits[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static __constant__ u32 kLLBase[36] = {0,  1,  2,  3,  4,  5,  6,  7,  8,  9,   10,  11,  12,   13,   14,   15,   16,    18,
                                20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
static __constant__ u8 kMLBits[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                               0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
// match length = kMLBase[code] + extra (already includes the minimum match of 3)
static __constant__ u32 kMLBase[53] = {3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 17, 18,  19,  20,  21,   22,   23,   24,   25,    26,    27,    28, 29,
                                30, 31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539};
// code of litLength <= 63 and of (matchLength - 3) <= 127
static __constant__ u8 kLLCode[64] = {0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 16, 17, 17, 18, 18, 19, 19, 20, 20, 20, 20, 21, 21, 21, 21,
                               22, 22, 22, 22, 22, 22, 22, 22, 23, 23, 23, 23, 23, 23, 23, 23, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24};
static __constant__ u8 kMLCode[128] = {0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31,
                                32, 32, 33, 33, 34, 34, 35, 35, 36, 36, 36, 36, 37, 37, 37, 37, 38, 38, 38, 38, 38, 38, 38, 38, 39, 39, 39, 39, 39, 39, 39, 39,
                                40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41,
                                42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42};
// predefined distributions (RFC 8878 §3.1.1.3.2.2)
static __constant__ short kLLDefNorm[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
static __constant__ short kMLDefNorm[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                     1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
static __constant__ short kOFDefNorm[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};

__device__ __forceinline__ u32 hibit32(u32 v) { return 31u - (u32)__clz((int)v); }  // v != 0

__device__ __forceinline__ u32 zstd_ll_code(u32 ll) { return ll > 63 ? hibit32(ll) + 19u : kLLCode[ll]; }
__device__ __forceinline__ u32 zstd_ml_code(u32 mlbase) { return mlbase > 127 ? hibit32(mlbase) + 36u : kMLCode[mlbase]; }

// XXH64 primes
#define XXP1 11400714785074694791ull
#define XXP2 14029467366897019727ull
#define XXP3 1609587929392839161ull
#define XXP4 9650029242287828579ull
#define XXP5 2870177450012600261ull
