/*
 * silo_pdb.c -- see silo_pdb.h.  C99, no dependencies.
 *
 * File layout (PDB format II, little-endian x86-64 data standard):
 *
 *   "!<<PDB:II>>!\n"
 *   primitive formats: one length byte n, then n-1 bytes: sizes of pointer, short, int, long, float,
 *       double; byte-order flags of short, int, long (2 = least significant byte first); the byte
 *       permutations of float and double; seven format fields of float and of double (bits, exponent
 *       bits, mantissa bits, sign position, exponent position, mantissa position, implicit-one flag)
 *   "<float bias>\001<double bias>\001\n"
 *   128 bytes reserved; at close they receive "<chart address>\001<symbol table address>\001\n"
 *   data of the variables, back to back
 *   structure chart: "type\001size\001[member\001...]\n" per type, closed by "\002\n"
 *   symbol table:    "name\001type\001nitems\001address\001[min\001count\001 per dimension]\n" per
 *                    variable, closed by an empty line
 *   extras:          Offset, Alignment, Struct-Alignment, Casts, Blocks, Major-Order, Has-Directories,
 *                    Version
 *
 * A Silo object is a variable of the struct type "Group" { char *name; char *type; char **comp_names;
 * char **pdb_names; integer ncomponents; }.  PDB stores what a pointer refers to right behind the
 * struct, every pointee introduced by an "itag" line "nitems\001type\001address\001flag\001\n"
 * (flag 1 = the data follows here; nitems 0 and address -1 = NULL).  pdb_names[i] is either the name of
 * a PDB variable holding component i ("/mesh_coord0") or a literal: '<i>3', '<d>0.5', '<s>text'.
 */
#define _FILE_OFFSET_BITS 64
#define _POSIX_C_SOURCE 200809L
#include "silo_pdb.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define SPDB_MAX_COMP 48

typedef struct spdb_sym {
    char name[128];
    char type[16];
    long long nitems, addr;
} spdb_sym;

typedef struct spdb_comps {
    int n;
    char names[SPDB_MAX_COMP][24];
    char values[SPDB_MAX_COMP][160];
} spdb_comps;

struct spdb_file {
    FILE *fp;
    long long headaddr;
    spdb_sym *syms;
    int nsyms, cap;
    int error;
    /* quadvar being streamed */
    char var_name[64], var_mesh[64];
    int var_dims[3];
    long long var_total, var_done;
};

static long long tell(spdb_file *f) { return (long long)ftello(f->fp); }

static void put_bytes(spdb_file *f, const void *p, size_t n)
{
    if (!f->error && n && fwrite(p, 1, n, f->fp) != n)
        f->error = 1;
}

static void put_text(spdb_file *f, const char *s) { put_bytes(f, s, strlen(s)); }

static spdb_sym *add_sym(spdb_file *f, const char *name, const char *type, long long nitems)
{
    if (f->nsyms == f->cap) {
        const int cap = f->cap ? 2 * f->cap : 64;
        spdb_sym *s = (spdb_sym *)realloc(f->syms, (size_t)cap * sizeof *s);
        if (!s) {
            f->error = 1;
            return NULL;
        }
        f->syms = s;
        f->cap = cap;
    }
    spdb_sym *s = &f->syms[f->nsyms++];
    snprintf(s->name, sizeof s->name, "%s", name);
    snprintf(s->type, sizeof s->type, "%s", type);
    s->nitems = nitems;
    s->addr = tell(f);
    return s;
}

/* one array variable, written in one go */
static void put_array(spdb_file *f, const char *name, const char *type, size_t elsize, const void *data, long long nitems)
{
    if (add_sym(f, name, type, nitems))
        put_bytes(f, data, elsize * (size_t)nitems);
}

static void put_string_var(spdb_file *f, const char *name, const char *text)
{
    put_array(f, name, "char", 1, text, (long long)strlen(text) + 1);
}

static void itag(spdb_file *f, long long nitems, const char *type)
{
    char line[96];
    /* flag 1: the data follows the tag; the address is that of the tag itself */
    snprintf(line, sizeof line, "%lld\001%s\001%lld\001%d\001\n", nitems, type, tell(f), 1);
    put_text(f, line);
}

static void put_pointee_string(spdb_file *f, const char *s)
{
    itag(f, (long long)strlen(s) + 1, "char");
    put_bytes(f, s, strlen(s) + 1);
}

static void put_pointee_string_array(spdb_file *f, int n, char (*strs)[160], char (*short_strs)[24])
{
    itag(f, n, "char *");
    for (int i = 0; i < n; ++i) { /* the pointer slots themselves: their values carry no meaning */
        const long long slot = i + 1;
        put_bytes(f, &slot, 8);
    }
    for (int i = 0; i < n; ++i)
        put_pointee_string(f, strs ? strs[i] : short_strs[i]);
}

/* <-> PJ_put_group: the object `name` of Silo type `type` with its components */
static void put_group(spdb_file *f, const char *name, const char *type, spdb_comps *c)
{
    char path[128];
    snprintf(path, sizeof path, "/%s", name);
    if (!add_sym(f, path, "Group", 1))
        return;
    unsigned char body[40];
    memset(body, 0, sizeof body);
    for (int i = 0; i < 4; ++i) {
        const long long slot = i + 1;
        memcpy(body + 8 * i, &slot, 8);
    }
    const int n = c->n;
    memcpy(body + 32, &n, 4);
    put_bytes(f, body, sizeof body);
    put_pointee_string(f, name);
    put_pointee_string(f, type);
    put_pointee_string_array(f, n, NULL, c->names);
    put_pointee_string_array(f, n, c->values, NULL);
}

static void comp_var(spdb_comps *c, const char *comp, const char *object, const char *suffix)
{
    if (c->n >= SPDB_MAX_COMP)
        return;
    snprintf(c->names[c->n], sizeof c->names[0], "%s", comp);
    snprintf(c->values[c->n], sizeof c->values[0], "/%s_%s", object, suffix);
    c->n++;
}

static void comp_int(spdb_comps *c, const char *comp, long long v)
{
    if (c->n >= SPDB_MAX_COMP)
        return;
    snprintf(c->names[c->n], sizeof c->names[0], "%s", comp);
    snprintf(c->values[c->n], sizeof c->values[0], "'<i>%lld'", v);
    c->n++;
}

static void comp_str(spdb_comps *c, const char *comp, const char *v)
{
    if (c->n >= SPDB_MAX_COMP)
        return;
    snprintf(c->names[c->n], sizeof c->names[0], "%s", comp);
    snprintf(c->values[c->n], sizeof c->values[0], "'<s>%s'", v);
    c->n++;
}

/* an array component: variable "/<object>_<suffix>" plus the entry in the component list */
static void comp_array(spdb_file *f, spdb_comps *c, const char *comp, const char *object, const char *type,
                       size_t elsize, const void *data, long long nitems)
{
    char path[128];
    snprintf(path, sizeof path, "/%s_%s", object, comp);
    put_array(f, path, type, elsize, data, nitems);
    comp_var(c, comp, object, comp);
}

enum { DB_COLLINEAR = 130, DB_RECTILINEAR = 100, DB_ROWMAJOR = 0, DB_OTHER = 124, DB_VOLUME = 141, DB_FLOAT = 19,
       DB_DOUBLE = 20, DB_ZONECENT = 111, DB_QUAD_RECT = 130, DB_QUADVAR = 501 };

spdb_file *spdb_create(const char *path, const char *fileinfo)
{
    FILE *fp = fopen(path, "wb");
    if (!fp)
        return NULL;
    spdb_file *f = (spdb_file *)calloc(1, sizeof *f);
    if (!f) {
        fclose(fp);
        return NULL;
    }
    f->fp = fp;
    put_text(f, "!<<PDB:II>>!\n");
    unsigned char fmt[64];
    int n = 1;
    const unsigned char sizes[6] = {8, 2, 4, 8, 4, 8};   /* pointer, short, int, long, float, double */
    memcpy(fmt + n, sizes, 6); n += 6;
    const unsigned char orders[3] = {2, 2, 2};          /* short, int, long: reverse (little-endian) order */
    memcpy(fmt + n, orders, 3); n += 3;
    for (int b = 4; b >= 1; --b) fmt[n++] = (unsigned char)b; /* float bytes, most significant first */
    for (int b = 8; b >= 1; --b) fmt[n++] = (unsigned char)b;
    const unsigned char ffmt[7] = {32, 8, 23, 0, 1, 9, 0}, dfmt[7] = {64, 11, 52, 0, 1, 12, 0};
    memcpy(fmt + n, ffmt, 7); n += 7;
    memcpy(fmt + n, dfmt, 7); n += 7;
    fmt[0] = (unsigned char)n;
    put_bytes(f, fmt, (size_t)n);
    put_text(f, "127\0011023\001\n");
    f->headaddr = tell(f);
    char pad[128];
    memset(pad, 0, sizeof pad);
    put_bytes(f, pad, sizeof pad);
    /* what DBCreate leaves in a new file: the root directory and three strings */
    const char zero = 0;
    put_array(f, "/", "Directory", 1, &zero, 1);
    put_string_var(f, "/_whatami", "linux-x86_64");
    put_string_var(f, "/_fileinfo", fileinfo ? fileinfo : "");
    put_string_var(f, "/_silolibinfo", "fdtd_b200 minimal PDB writer (Silo PDB-driver layout)");
    if (f->error) {
        spdb_close(f);
        return NULL;
    }
    return f;
}

int spdb_put_quadmesh(spdb_file *f, const char *name, const double *const coords[3], const int dims[3])
{
    static const char *const cn[3] = {"coord0", "coord1", "coord2"};
    spdb_comps c;
    memset(&c, 0, sizeof c);
    double lo[3], hi[3];
    int zero3[3] = {0, 0, 0}, max_index[3];
    long long nnodes = 1;
    for (int d = 0; d < 3; ++d) {
        comp_array(f, &c, cn[d], name, "double", 8, coords[d], dims[d]);
        lo[d] = coords[d][0];
        hi[d] = coords[d][dims[d] - 1];
        max_index[d] = dims[d] - 1;
        nnodes *= dims[d];
    }
    comp_array(f, &c, "min_extents", name, "double", 8, lo, 3);
    comp_array(f, &c, "max_extents", name, "double", 8, hi, 3);
    comp_int(&c, "ndims", 3);
    comp_int(&c, "coordtype", DB_COLLINEAR);
    comp_int(&c, "nspace", 3);
    comp_int(&c, "nnodes", nnodes);
    comp_int(&c, "facetype", DB_RECTILINEAR);
    comp_int(&c, "major_order", DB_ROWMAJOR);
    comp_int(&c, "cycle", 0);
    comp_int(&c, "coord_sys", DB_OTHER);
    comp_int(&c, "planar", DB_VOLUME);
    comp_int(&c, "origin", 0);
    comp_int(&c, "datatype", DB_DOUBLE);
    comp_array(f, &c, "dims", name, "integer", 4, dims, 3);
    comp_array(f, &c, "min_index", name, "integer", 4, zero3, 3);
    comp_array(f, &c, "max_index", name, "integer", 4, max_index, 3);
    comp_array(f, &c, "baseindex", name, "integer", 4, zero3, 3);
    put_group(f, name, "quadmesh", &c);
    return f->error ? -1 : 0;
}

int spdb_quadvar_begin(spdb_file *f, const char *name, const char *meshname, const int zdims[3])
{
    char path[128];
    snprintf(f->var_name, sizeof f->var_name, "%s", name);
    snprintf(f->var_mesh, sizeof f->var_mesh, "%s", meshname);
    memcpy(f->var_dims, zdims, sizeof f->var_dims);
    f->var_total = (long long)zdims[0] * zdims[1] * zdims[2];
    f->var_done = 0;
    snprintf(path, sizeof path, "/%s_data", name);
    add_sym(f, path, "double", f->var_total);
    return f->error ? -1 : 0;
}

int spdb_quadvar_append(spdb_file *f, const double *data, size_t count)
{
    if (f->var_done + (long long)count > f->var_total)
        f->error = 1;
    put_bytes(f, data, count * sizeof(double));
    f->var_done += (long long)count;
    return f->error ? -1 : 0;
}

int spdb_quadvar_end(spdb_file *f)
{
    if (f->var_done != f->var_total)
        f->error = 1;
    const char *name = f->var_name;
    spdb_comps c;
    memset(&c, 0, sizeof c);
    int zero3[3] = {0, 0, 0}, max_index[3];
    const float align[3] = {0.5f, 0.5f, 0.5f}; /* zone centred */
    for (int d = 0; d < 3; ++d)
        max_index[d] = f->var_dims[d] - 1;
    comp_str(&c, "meshid", f->var_mesh);
    comp_var(&c, "value0", name, "data");
    comp_int(&c, "ndims", 3);
    comp_int(&c, "nvals", 1);
    comp_int(&c, "nels", f->var_total);
    comp_int(&c, "origin", 0);
    comp_int(&c, "datatype", DB_DOUBLE);
    comp_int(&c, "centering", DB_ZONECENT);
    comp_int(&c, "mixlen", 0);
    comp_int(&c, "major_order", DB_ROWMAJOR);
    comp_int(&c, "cycle", 0);
    comp_array(f, &c, "dims", name, "integer", 4, f->var_dims, 3);
    comp_array(f, &c, "zones", name, "integer", 4, f->var_dims, 3);
    comp_array(f, &c, "min_index", name, "integer", 4, zero3, 3);
    comp_array(f, &c, "max_index", name, "integer", 4, max_index, 3);
    comp_array(f, &c, "align", name, "float", 4, align, 3);
    put_group(f, name, "quadvar", &c);
    return f->error ? -1 : 0;
}

/* Silo's string-list encoding: every string preceded by ';' */
static char *join_list(int n, const char *const strs[])
{
    size_t len = 1;
    for (int i = 0; i < n; ++i)
        len += strlen(strs[i]) + 1;
    char *out = (char *)malloc(len);
    if (!out)
        return NULL;
    out[0] = 0;
    for (int i = 0; i < n; ++i) {
        strcat(out, ";");
        strcat(out, strs[i]);
    }
    return out;
}

int spdb_put_defvars(spdb_file *f, const char *name, int n, const char *const names[], const int types[],
                     const char *const defns[])
{
    spdb_comps c;
    memset(&c, 0, sizeof c);
    char *nl = join_list(n, names), *dl = join_list(n, defns);
    if (!nl || !dl) {
        free(nl);
        free(dl);
        f->error = 1;
        return -1;
    }
    comp_int(&c, "ndefs", n);
    comp_array(f, &c, "names", name, "char", 1, nl, (long long)strlen(nl) + 1);
    comp_array(f, &c, "types", name, "integer", 4, types, n);
    comp_array(f, &c, "defns", name, "char", 1, dl, (long long)strlen(dl) + 1);
    put_group(f, name, "defvars", &c);
    free(nl);
    free(dl);
    return f->error ? -1 : 0;
}

static int put_multi(spdb_file *f, const char *name, const char *type, const char *names_comp, const char *types_comp,
                     int block_type, int nblocks, const char *const blocknames[])
{
    spdb_comps c;
    memset(&c, 0, sizeof c);
    char *nl = join_list(nblocks, blocknames);
    int *types = (int *)malloc(sizeof(int) * (size_t)(nblocks > 0 ? nblocks : 1));
    if (!nl || !types) {
        free(nl);
        free(types);
        f->error = 1;
        return -1;
    }
    for (int i = 0; i < nblocks; ++i)
        types[i] = block_type;
    comp_int(&c, "nblocks", nblocks);
    comp_int(&c, "ngroups", 0);
    comp_int(&c, "blockorigin", 0);
    comp_int(&c, "grouporigin", 0);
    comp_array(f, &c, types_comp, name, "integer", 4, types, nblocks);
    comp_array(f, &c, names_comp, name, "char", 1, nl, (long long)strlen(nl) + 1);
    put_group(f, name, type, &c);
    free(nl);
    free(types);
    return f->error ? -1 : 0;
}

int spdb_put_multimesh(spdb_file *f, const char *name, int nblocks, const char *const blocknames[])
{
    return put_multi(f, name, "multimesh", "meshnames", "meshtypes", DB_QUAD_RECT, nblocks, blocknames);
}

int spdb_put_multivar(spdb_file *f, const char *name, int nblocks, const char *const blocknames[])
{
    return put_multi(f, name, "multivar", "varnames", "vartypes", DB_QUADVAR, nblocks, blocknames);
}

int spdb_close(spdb_file *f)
{
    if (!f)
        return 0;
    const long long chart = tell(f);
    put_text(f, "*\0018\001\n"
                "short\0012\001\n"
                "integer\0014\001\n"
                "int\0014\001\n"
                "long\0018\001\n"
                "float\0014\001\n"
                "double\0018\001\n"
                "char\0011\001\n"
                "Directory\0011\001\n"
                "Group\00140\001char *name\001char *type\001char **comp_names\001char **pdb_names\001integer ncomponents\001\n"
                "\002\n");
    const long long symtab = tell(f);
    for (int i = 0; i < f->nsyms; ++i) {
        char line[320];
        const spdb_sym *s = &f->syms[i];
        if (!strcmp(s->type, "Group") || !strcmp(s->type, "Directory")) /* scalars carry no dimensions */
            snprintf(line, sizeof line, "%s\001%s\001%lld\001%lld\001\n", s->name, s->type, s->nitems, s->addr);
        else
            snprintf(line, sizeof line, "%s\001%s\001%lld\001%lld\0010\001%lld\001\n", s->name, s->type, s->nitems,
                     s->addr, s->nitems);
        put_text(f, line);
    }
    put_text(f, "\n");
    /* extras: alignments of char, pointer, short, int, long, float, double as raw bytes */
    put_text(f, "Offset:0\n");
    put_text(f, "Alignment:\001\010\002\004\010\004\010\n");
    put_text(f, "Struct-Alignment:0\n");
    put_text(f, "Casts:\n\002\n");
    put_text(f, "Blocks:\n\002\n");
    put_text(f, "Major-Order:101\n");
    put_text(f, "Has-Directories:1\n");
    put_text(f, "Version:14|Sun Oct 18 00:00:00 2026\n");
    put_text(f, "\n\n");
    if (!f->error && fseeko(f->fp, (off_t)f->headaddr, SEEK_SET) == 0) {
        char line[64];
        snprintf(line, sizeof line, "%lld\001%lld\001\n", chart, symtab);
        put_text(f, line);
    } else {
        f->error = 1;
    }
    int rc = f->error ? -1 : 0;
    if (fclose(f->fp) != 0)
        rc = -1;
    free(f->syms);
    free(f);
    return rc;
}
