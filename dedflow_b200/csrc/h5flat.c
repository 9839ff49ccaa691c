/* h5flat.c -- DEDFlow's H5* file interface over a flat container file (include/dedflow_h5flat.h): what the reference's
 * Mesh3DCreateH5 (src/Mesh.c:12-104), Mesh3DDataCreateH5 (src/MeshData.c:57-109) and main() (src/main.c:366-368,521-532,571-591)
 * call, for machines without libhdf5.  Plain C99, no dependencies; one open-file table, not thread safe (like the reference). */
#include "../../include/dedflow_h5flat.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const char kMagic[8] = {'D', 'F', 'B', 'H', '5', 0, 1, 0};
enum { DT_I32 = 0, DT_U32 = 1, DT_F32 = 2, DT_F64 = 3, DT_I64 = 4, DT_U64 = 5 };
static const size_t kSize[6] = {4, 4, 4, 8, 8, 8};

typedef struct {
  char* name;
  int dtype;
  uint64_t count;
  long offset; /* of the data in the file */
} Entry;

typedef struct {
  FILE* fp;
  int writable, readable;
  Entry* ent;
  size_t n_ent, cap;
} Flat;

#define MAX_OPEN 64
static Flat* g_open[MAX_OPEN];

static void die(const char* what, const char* arg) {
  /* the reference ASSERTs (traps) on a failed open or read (src/h5util.c:21); say why first */
  fprintf(stderr, "dedflow h5flat: %s: %s\n", what, arg ? arg : "");
  abort();
}

static Flat* flat_of(H5FileInfo* f) {
  if (!f || f->file_id < 1 || f->file_id > MAX_OPEN || !g_open[f->file_id - 1]) die("invalid file handle", f ? f->filename : "(null)");
  return g_open[f->file_id - 1];
}

static void add_entry(Flat* F, const char* name, size_t len, int dtype, uint64_t count, long offset) {
  for (size_t i = 0; i < F->n_ent; i++)
    if (strlen(F->ent[i].name) == len && !memcmp(F->ent[i].name, name, len)) {   /* replaced by the later record */
      F->ent[i].dtype = dtype; F->ent[i].count = count; F->ent[i].offset = offset;
      return;
    }
  if (F->n_ent == F->cap) {
    F->cap = F->cap ? 2 * F->cap : 32;
    F->ent = (Entry*)realloc(F->ent, F->cap * sizeof(Entry));
    if (!F->ent) die("out of memory", NULL);
  }
  Entry* e = &F->ent[F->n_ent++];
  e->name = (char*)malloc(len + 1);
  memcpy(e->name, name, len);
  e->name[len] = 0;
  e->dtype = dtype; e->count = count; e->offset = offset;
}

/* returns 0 when the file is not a container */
static int scan(Flat* F) {
  char magic[8];
  if (fseek(F->fp, 0, SEEK_SET) || fread(magic, 1, 8, F->fp) != 8 || memcmp(magic, kMagic, 8)) return 0;
  for (;;) {
    uint32_t nl;
    if (fread(&nl, 4, 1, F->fp) != 1) break;   /* clean end of file */
    if (nl == 0 || nl > 4096) return 0;
    char name[4097];
    uint8_t dt;
    uint64_t cnt;
    if (fread(name, 1, nl, F->fp) != nl || fread(&dt, 1, 1, F->fp) != 1 || fread(&cnt, 8, 1, F->fp) != 1 || dt > DT_U64) return 0;
    const long off = ftell(F->fp);
    add_entry(F, name, nl, dt, cnt, off);
    if (fseek(F->fp, (long)(cnt * kSize[dt]), SEEK_CUR)) return 0;
  }
  return 1;
}

static const char* strip(const char* name) {
  while (*name == '/') name++;
  return name;
}

static Entry* find(Flat* F, const char* name) {
  name = strip(name);
  for (size_t i = 0; i < F->n_ent; i++)
    if (!strcmp(F->ent[i].name, name)) return &F->ent[i];
  return NULL;
}

H5FileInfo* H5OpenFile(const char* filename, const char* mode) {
  if (!filename || !mode) die("H5OpenFile: bad argument", filename);
  int slot = -1;
  for (int i = 0; i < MAX_OPEN; i++)
    if (!g_open[i]) { slot = i; break; }
  if (slot < 0) die("H5OpenFile: too many open files", filename);
  Flat* F = (Flat*)calloc(1, sizeof(Flat));
  if (!strcmp(mode, "r")) {
    F->fp = fopen(filename, "rb");
    F->readable = 1;
  } else if (!strcmp(mode, "w")) {
    F->fp = fopen(filename, "w+b");
    F->writable = 1;
    if (F->fp) fwrite(kMagic, 1, 8, F->fp);
  } else if (!strcmp(mode, "a")) {
    F->fp = fopen(filename, "r+b");
    F->readable = F->writable = 1;
  } else {
    die("H5OpenFile: invalid mode", mode);
  }
  if (!F->fp) die("H5OpenFile: failed to open file", filename);
  if (strcmp(mode, "w") && !scan(F)) die("H5OpenFile: not a dedflow flat container (convert it with dedflow_b200/h5flat.py)", filename);
  g_open[slot] = F;
  H5FileInfo* h = (H5FileInfo*)calloc(1, sizeof(H5FileInfo));
  strncpy(h->filename, filename, sizeof(h->filename) - 1);
  h->file_id = slot + 1;
  return h;
}

void H5CloseFile(H5FileInfo* h) {
  Flat* F = flat_of(h);
  fclose(F->fp);
  for (size_t i = 0; i < F->n_ent; i++) free(F->ent[i].name);
  free(F->ent);
  free(F);
  g_open[h->file_id - 1] = NULL;
  free(h);
}

int32_t H5FileExist(const char* filename) {
  FILE* fp = fopen(filename, "rb");
  if (!fp) return 0;
  char magic[8];
  const int ok = fread(magic, 1, 8, fp) == 8 && !memcmp(magic, kMagic, 8);
  fclose(fp);
  return ok;
}

int32_t H5FileIsWritable(H5FileInfo* h) { return flat_of(h)->writable; }
int32_t H5FileIsReadable(H5FileInfo* h) { return flat_of(h)->readable; }

int32_t H5GroupExist(H5FileInfo* h, const char* group_name) {
  Flat* F = flat_of(h);
  const char* g = strip(group_name);
  size_t len = strlen(g);
  while (len && g[len - 1] == '/') len--;
  if (!len) return 1;   /* the root group */
  for (size_t i = 0; i < F->n_ent; i++)
    if (!strncmp(F->ent[i].name, g, len) && F->ent[i].name[len] == '/') return 1;
  return 0;
}

int32_t H5DatasetExist(H5FileInfo* h, const char* dataset_name) { return find(flat_of(h), dataset_name) != NULL; }

void H5GetDatasetSize(H5FileInfo* h, const char* dataset_name, int32_t* size) {
  Entry* e = find(flat_of(h), dataset_name);
  if (!e) { *size = 0; return; }
  if (e->count >= 0x7fffffffull) die("H5GetDatasetSize: dataset exceeds the limit of index_type", dataset_name);
  *size = (int32_t)e->count;
}

static double load_as_f64(const unsigned char* p, int dt) {
  switch (dt) {
    case DT_I32: { int32_t v; memcpy(&v, p, 4); return (double)v; }
    case DT_U32: { uint32_t v; memcpy(&v, p, 4); return (double)v; }
    case DT_F32: { float v; memcpy(&v, p, 4); return (double)v; }
    case DT_F64: { double v; memcpy(&v, p, 8); return v; }
    case DT_I64: { int64_t v; memcpy(&v, p, 8); return (double)v; }
    default: { uint64_t v; memcpy(&v, p, 8); return (double)v; }
  }
}
static int64_t load_as_i64(const unsigned char* p, int dt) {
  switch (dt) {
    case DT_I32: { int32_t v; memcpy(&v, p, 4); return v; }
    case DT_U32: { uint32_t v; memcpy(&v, p, 4); return v; }
    case DT_F32: { float v; memcpy(&v, p, 4); return (int64_t)v; }
    case DT_F64: { double v; memcpy(&v, p, 8); return (int64_t)v; }
    case DT_I64: { int64_t v; memcpy(&v, p, 8); return v; }
    default: { uint64_t v; memcpy(&v, p, 8); return (int64_t)v; }
  }
}

/* read the whole dataset into `out` (type `want`), converting when the stored type differs */
static void read_dataset(H5FileInfo* h, const char* name, int want, void* out) {
  Flat* F = flat_of(h);
  Entry* e = find(F, name);
  if (!e) die("H5ReadDataset: no such dataset", name);
  if (fseek(F->fp, e->offset, SEEK_SET)) die("H5ReadDataset: seek failed", name);
  if (e->dtype == want) {
    if (fread(out, kSize[want], e->count, F->fp) != e->count) die("H5ReadDataset: short read", name);
    return;
  }
  enum { CHUNK = 8192 };
  unsigned char buf[CHUNK * 8];
  const size_t es = kSize[e->dtype];
  uint64_t done = 0;
  while (done < e->count) {
    const size_t n = e->count - done < CHUNK ? (size_t)(e->count - done) : CHUNK;
    if (fread(buf, es, n, F->fp) != n) die("H5ReadDataset: short read", name);
    for (size_t i = 0; i < n; i++) {
      const unsigned char* p = buf + i * es;
      switch (want) {
        case DT_I32: ((int32_t*)out)[done + i] = (int32_t)load_as_i64(p, e->dtype); break;
        case DT_U32: ((uint32_t*)out)[done + i] = (uint32_t)load_as_i64(p, e->dtype); break;
        case DT_F32: ((float*)out)[done + i] = (float)load_as_f64(p, e->dtype); break;
        default: ((double*)out)[done + i] = load_as_f64(p, e->dtype); break;
      }
    }
    done += n;
  }
}

static void write_dataset(H5FileInfo* h, const char* name, int dt, int32_t len, const void* data) {
  Flat* F = flat_of(h);
  if (!F->writable) die("H5WriteDataset: file is not writable", h->filename);
  if (len < 0) die("H5WriteDataset: negative length", name);
  name = strip(name);
  const uint32_t nl = (uint32_t)strlen(name);
  const uint8_t d8 = (uint8_t)dt;
  const uint64_t cnt = (uint64_t)len;
  if (fseek(F->fp, 0, SEEK_END)) die("H5WriteDataset: seek failed", name);
  if (fwrite(&nl, 4, 1, F->fp) != 1 || fwrite(name, 1, nl, F->fp) != nl || fwrite(&d8, 1, 1, F->fp) != 1 || fwrite(&cnt, 8, 1, F->fp) != 1)
    die("H5WriteDataset: write failed", name);
  const long off = ftell(F->fp);
  if (cnt && fwrite(data, kSize[dt], cnt, F->fp) != cnt) die("H5WriteDataset: write failed", name);
  fflush(F->fp);
  add_entry(F, name, nl, dt, cnt, off);
}

void H5ReadDataseti32(H5FileInfo* h, const char* n, int32_t* d) { read_dataset(h, n, DT_I32, d); }
void H5ReadDatasetu32(H5FileInfo* h, const char* n, uint32_t* d) { read_dataset(h, n, DT_U32, d); }
void H5ReadDatasetf32(H5FileInfo* h, const char* n, float* d) { read_dataset(h, n, DT_F32, d); }
void H5ReadDatasetf64(H5FileInfo* h, const char* n, double* d) { read_dataset(h, n, DT_F64, d); }
void H5ReadDatasetInd(H5FileInfo* h, const char* n, int32_t* d) { read_dataset(h, n, DT_I32, d); }
void H5ReadDatasetVal(H5FileInfo* h, const char* n, double* d) { read_dataset(h, n, DT_F64, d); }
void H5WriteDataseti32(H5FileInfo* h, const char* n, int32_t len, const int32_t* d) { write_dataset(h, n, DT_I32, len, d); }
void H5WriteDatasetu32(H5FileInfo* h, const char* n, int32_t len, const uint32_t* d) { write_dataset(h, n, DT_U32, len, d); }
void H5WriteDatasetf32(H5FileInfo* h, const char* n, int32_t len, const float* d) { write_dataset(h, n, DT_F32, len, d); }
void H5WriteDatasetf64(H5FileInfo* h, const char* n, int32_t len, const double* d) { write_dataset(h, n, DT_F64, len, d); }
void H5WriteDatasetInd(H5FileInfo* h, const char* n, int32_t len, const int32_t* d) { write_dataset(h, n, DT_I32, len, d); }
void H5WriteDatasetVal(H5FileInfo* h, const char* n, int32_t len, const double* d) { write_dataset(h, n, DT_F64, len, d); }
