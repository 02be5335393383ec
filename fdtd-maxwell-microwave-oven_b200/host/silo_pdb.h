/*
 * silo_pdb.h -- minimal writer of Silo files in the PDB driver's on-disk format, for the objects
 * write_silo() of the reference creates (main.c:550-598): one collinear quadmesh, zone-centred double
 * quadvars on it, one defvars object -- plus multimesh / multivar objects for runs split into z-slabs.
 *
 * libsilo is not available in this image ("Facts at a glance"), so the files are produced
 * directly: a PDB (format II) container -- ASCII header, primitive-format block, data, structure
 * chart, symbol table, extras -- holding the variables and the "Group" objects Silo's PDB driver
 * writes for DBPutQuadmesh / DBPutQuadvar1 / DBPutDefvars.  The layout follows the published PDBLib
 * file format and Silo's silo_pdb.c naming (<object>_<component> arrays, a Group per object).  It has
 * NOT been cross-checked against libsilo itself (none here); tests/pdb_reader.py is an independent
 * reader of the container and of the Group encoding.
 */
#ifndef SILO_PDB_H
#define SILO_PDB_H

#include <stddef.h>

typedef struct spdb_file spdb_file;

/* <-> DBCreate(path, DB_CLOBBER, DB_LOCAL, fileinfo, DB_PDB), main.c:555.  NULL on failure. */
spdb_file *spdb_create(const char *path, const char *fileinfo);
/* <-> DBPutQuadmesh(db, name, NULL, coords, dims, 3, DB_DOUBLE, DB_COLLINEAR, NULL), main.c:561 */
int spdb_put_quadmesh(spdb_file *f, const char *name, const double *const coords[3], const int dims[3]);
/* <-> DBPutQuadvar1(db, name, mesh, data, zdims, 3, NULL, 0, DB_DOUBLE, DB_ZONECENT, NULL), main.c:564.
 * The data may be handed over in pieces: begin (declares the total), any number of appends, end. */
int spdb_quadvar_begin(spdb_file *f, const char *name, const char *meshname, const int zdims[3]);
int spdb_quadvar_append(spdb_file *f, const double *data, size_t count);
int spdb_quadvar_end(spdb_file *f);
/* <-> DBPutDefvars(db, name, n, names, types, defns, NULL), main.c:595 */
int spdb_put_defvars(spdb_file *f, const char *name, int n, const char *const names[], const int types[],
                     const char *const defns[]);
/* <-> DBPutMultimesh / DBPutMultivar: block names are "file:/object" */
int spdb_put_multimesh(spdb_file *f, const char *name, int nblocks, const char *const blocknames[]);
int spdb_put_multivar(spdb_file *f, const char *name, int nblocks, const char *const blocknames[]);
/* <-> DBClose, main.c:597: structure chart, symbol table, extras, header addresses */
int spdb_close(spdb_file *f);

/* constants of silo.h used above */
enum { SPDB_VARTYPE_SCALAR = 200, SPDB_VARTYPE_VECTOR = 201 };

#endif
