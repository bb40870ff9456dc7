/* oracle/tbin.h -- TEST INFRASTRUCTURE.  Tiny named-array container used to hand identical inputs to
 * the reference harness (oracle/_ref/taco_ref_harness), the C oracle and the Python tests.
 *
 *   file   := "TB2BIN\0\0" u32 n_arrays  u32 0   array*
 *   array  := char name[24] (NUL padded)  u32 dtype  u32 0  u64 count  payload (count*size, padded to 8 B)
 *   dtype  := 0 int32 | 1 float32 | 2 float64 | 3 int64
 *
 * The Python twin lives in taco_b200/tbin.py.
 */
#ifndef TACO_B200_TBIN_H
#define TACO_B200_TBIN_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { char name[24]; uint32_t dtype; uint64_t count; void* data; } tbin_array;
typedef struct { uint32_t n; tbin_array* a; } tbin_file;

static inline size_t tbin_esize(uint32_t dtype) { return dtype == 0 ? 4 : dtype == 1 ? 4 : 8; }

static inline int tbin_read(const char* path, tbin_file* out) {
  FILE* f = fopen(path, "rb");
  if (!f) return -1;
  char magic[8]; uint32_t n, z;
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "TB2BIN\0\0", 8) != 0) { fclose(f); return -2; }
  if (fread(&n, 4, 1, f) != 1 || fread(&z, 4, 1, f) != 1) { fclose(f); return -2; }
  out->n = n; out->a = (tbin_array*)calloc(n ? n : 1, sizeof(tbin_array));
  for (uint32_t i = 0; i < n; i++) {
    tbin_array* a = &out->a[i];
    if (fread(a->name, 1, 24, f) != 24) { fclose(f); return -3; }
    if (fread(&a->dtype, 4, 1, f) != 1 || fread(&z, 4, 1, f) != 1 || fread(&a->count, 8, 1, f) != 1) { fclose(f); return -3; }
    size_t bytes = a->count * tbin_esize(a->dtype), padded = (bytes + 7) & ~(size_t)7;
    a->data = malloc(padded ? padded : 8);
    if (padded && fread(a->data, 1, padded, f) != padded) { fclose(f); return -4; }
  }
  fclose(f);
  return 0;
}

static inline tbin_array* tbin_get(tbin_file* tf, const char* name) {
  for (uint32_t i = 0; i < tf->n; i++) if (strncmp(tf->a[i].name, name, 24) == 0) return &tf->a[i];
  return NULL;
}

static inline int tbin_write(const char* path, const tbin_array* arrs, uint32_t n) {
  FILE* f = fopen(path, "wb");
  if (!f) return -1;
  uint32_t z = 0;
  fwrite("TB2BIN\0\0", 1, 8, f); fwrite(&n, 4, 1, f); fwrite(&z, 4, 1, f);
  for (uint32_t i = 0; i < n; i++) {
    char name[24]; memset(name, 0, 24); strncpy(name, arrs[i].name, 23);
    fwrite(name, 1, 24, f); fwrite(&arrs[i].dtype, 4, 1, f); fwrite(&z, 4, 1, f); fwrite(&arrs[i].count, 8, 1, f);
    size_t bytes = arrs[i].count * tbin_esize(arrs[i].dtype), padded = (bytes + 7) & ~(size_t)7;
    fwrite(arrs[i].data, 1, bytes, f);
    static const char pad[8] = {0};
    if (padded > bytes) fwrite(pad, 1, padded - bytes, f);
  }
  fclose(f);
  return 0;
}
#endif
