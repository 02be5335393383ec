/*
 * silo.h -- stand-in for libsilo's header, used ONLY to compile the unmodified reference
 * (/root/reference/main.c:15 includes <silo.h>) into oracle/_ref/ for parity checks.
 * libsilo is not installed in this image and there is no network.  Test infrastructure.
 *
 * It declares exactly what main.c:555-597 uses.  Instead of writing a file, every call is
 * forwarded to an optional recorder so tests can see what the reference would have dumped.
 */
#ifndef ORACLE_SILO_STUB_H
#define ORACLE_SILO_STUB_H

#include <stddef.h>

typedef struct DBfile { int open; } DBfile;
typedef struct DBoptlist DBoptlist;

enum { DB_CLOBBER = 0, DB_LOCAL = 0, DB_PDB = 2, DB_DOUBLE = 20, DB_COLLINEAR = 130,
       DB_ZONECENT = 111, DB_VARTYPE_VECTOR = 201 };

/* recorder: kind 0 = create(name), 1 = quadvar(name, data, n = product of dims), 2 = close */
typedef void (*silo_stub_recorder)(int kind, const char *name, const double *data, size_t n);
static silo_stub_recorder silo_stub_hook;
static DBfile silo_stub_file;

static DBfile *DBCreate(const char *name, int mode, int target, const char *info, int type)
{
    (void)mode; (void)target; (void)info; (void)type;
    if (silo_stub_hook) silo_stub_hook(0, name, NULL, 0);
    silo_stub_file.open = 1;
    return &silo_stub_file;
}

static int DBPutQuadmesh(DBfile *f, const char *name, const char *const *coordnames, void *coords,
                         int *dims, int ndims, int datatype, int coordtype, DBoptlist *opts)
{
    (void)f; (void)name; (void)coordnames; (void)coords; (void)dims; (void)ndims;
    (void)datatype; (void)coordtype; (void)opts;
    return 0;
}

static int DBPutQuadvar1(DBfile *f, const char *name, const char *meshname, void *var, int *dims,
                         int ndims, void *mixvar, int mixlen, int datatype, int centering,
                         DBoptlist *opts)
{
    size_t n = 1;
    (void)f; (void)meshname; (void)mixvar; (void)mixlen; (void)datatype; (void)centering; (void)opts;
    for (int d = 0; d < ndims; ++d) n *= (size_t)dims[d];
    if (silo_stub_hook) silo_stub_hook(1, name, (const double *)var, n);
    return 0;
}

static int DBPutDefvars(DBfile *f, const char *name, int ndefs, const char *const *names,
                        const int *types, const char *const *defs, DBoptlist *const *opts)
{
    (void)f; (void)name; (void)ndefs; (void)names; (void)types; (void)defs; (void)opts;
    return 0;
}

static int DBClose(DBfile *f)
{
    f->open = 0;
    if (silo_stub_hook) silo_stub_hook(2, NULL, NULL, 0);
    return 0;
}

#endif
