/* petsc_mock.c -- behaviour of the PETSc slice declared in petsc_mock.h (TEST INFRASTRUCTURE; see the header). */
#include "petsc_mock.h"
#include <stdarg.h>

static char g_err[1024];
PetscErrorCode PetscMockError(int code, const char *fmt, ...)
{ va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap); return code ? code : 1; }
const char *PetscMockLastError(void) { return g_err; }

/* ---- Vec */
PetscErrorCode VecCreateSeq(MPI_Comm c, PetscInt n, Vec *v) { (void)c; *v = calloc(1, sizeof(**v)); (*v)->n = n; (*v)->a = calloc(n > 0 ? n : 1, sizeof(PetscScalar)); return 0; }
PetscErrorCode VecDestroy(Vec *v) { if (*v) { if ((*v)->rd || (*v)->wr) return PetscMockError(PETSC_ERR_ORDER, "VecDestroy: array not restored"); free((*v)->a); free(*v); *v = NULL; } return 0; }
PetscErrorCode VecGetSize(Vec v, PetscInt *n) { *n = v->n; return 0; }
PetscErrorCode VecGetArrayRead(Vec v, const PetscScalar **a) { v->rd++; *a = v->a; return 0; }
PetscErrorCode VecRestoreArrayRead(Vec v, const PetscScalar **a) { if (v->rd < 1) return PetscMockError(PETSC_ERR_ORDER, "VecRestoreArrayRead without Get"); v->rd--; *a = NULL; return 0; }
PetscErrorCode VecGetArray(Vec v, PetscScalar **a) { if (v->wr) return PetscMockError(PETSC_ERR_ORDER, "VecGetArray: already in use"); v->wr++; *a = v->a; return 0; }
PetscErrorCode VecRestoreArray(Vec v, PetscScalar **a) { if (v->wr < 1) return PetscMockError(PETSC_ERR_ORDER, "VecRestoreArray without Get"); v->wr--; *a = NULL; return 0; }

/* ---- registries */
#define MAXREG 16
static struct { char name[64]; PetscErrorCode (*create)(Mat); } g_mat[MAXREG]; static int g_nmat;
static struct { char name[64]; PetscErrorCode (*create)(PC); } g_pc[MAXREG]; static int g_npc;
PetscErrorCode MatRegister(const char *n, PetscErrorCode (*f)(Mat)) { if (g_nmat == MAXREG) return 1; snprintf(g_mat[g_nmat].name, 64, "%s", n); g_mat[g_nmat++].create = f; return 0; }
PetscErrorCode PCRegister(const char *n, PetscErrorCode (*f)(PC)) { if (g_npc == MAXREG) return 1; snprintf(g_pc[g_npc].name, 64, "%s", n); g_pc[g_npc++].create = f; return 0; }

/* ---- Mat */
PetscErrorCode MatCreate(MPI_Comm c, Mat *A) { (void)c; *A = calloc(1, sizeof(**A)); return 0; }
PetscErrorCode MatSetSizes(Mat A, PetscInt m, PetscInt n, PetscInt M, PetscInt N) { (void)M; (void)N; A->m = m; A->n = n; return 0; }
PetscErrorCode MatGetSize(Mat A, PetscInt *m, PetscInt *n) { if (m) *m = A->m; if (n) *n = A->n; return 0; }
PetscErrorCode MatSetType(Mat A, MatType t)
{
  if (!strcmp(t, MATSHELL)) { snprintf(A->type, 64, "%s", t); return 0; }
  for (int i = 0; i < g_nmat; ++i) if (!strcmp(g_mat[i].name, t)) { PetscErrorCode e = g_mat[i].create(A); if (!e) snprintf(A->type, 64, "%s", t); return e; }
  return PetscMockError(PETSC_ERR_ARG_WRONG, "Unknown Mat type %s", t);
}
PetscErrorCode MatShellSetContext(Mat A, void *ctx) { A->shellctx = ctx; return 0; }
PetscErrorCode MatShellGetContext(Mat A, void *ctx) { *(void **)ctx = A->shellctx; return 0; }
PetscErrorCode MatShellSetOperation(Mat A, MatOperation op, void (*f)(void))
{
  if (op == MATOP_MULT) A->mult = (PetscErrorCode(*)(Mat, Vec, Vec))f;
  else if (op == MATOP_GET_DIAGONAL) A->getdiagonal = (PetscErrorCode(*)(Mat, Vec))f;
  else if (op == MATOP_DESTROY) A->destroy = (PetscErrorCode(*)(Mat))f;
  else return PetscMockError(PETSC_ERR_SUP, "mock MatShellSetOperation: op %d", (int)op);
  return 0;
}
PetscErrorCode MatMult(Mat A, Vec x, Vec y) { if (!A->mult) return PetscMockError(PETSC_ERR_SUP, "No MatMult for type %s", A->type); if (x == y) return PetscMockError(PETSC_ERR_ARG_WRONG, "x and y must differ"); return A->mult(A, x, y); }
PetscErrorCode MatGetDiagonal(Mat A, Vec d) { if (!A->getdiagonal) return PetscMockError(PETSC_ERR_SUP, "No MatGetDiagonal for type %s", A->type); return A->getdiagonal(A, d); }
PetscErrorCode MatDestroy(Mat *A) { PetscErrorCode e = 0; if (*A) { if ((*A)->destroy) e = (*A)->destroy(*A); free(*A); *A = NULL; } return e; }
PetscErrorCode PetscObjectTypeCompare(PetscObject o, const char *t, PetscBool *same) { (void)o; *same = (!strcmp(t, PETSCVIEWERASCII)) ? PETSC_TRUE : PETSC_FALSE; return 0; }

/* ---- options database: one string of "-key value" tokens */
static char *g_opts;
PetscErrorCode PetscOptionsClear(void *o) { (void)o; free(g_opts); g_opts = NULL; return 0; }
PetscErrorCode PetscOptionsInsertString(void *o, const char *s)
{ (void)o; size_t a = g_opts ? strlen(g_opts) : 0, b = strlen(s); g_opts = realloc(g_opts, a + b + 2); if (a) g_opts[a++] = ' '; memcpy(g_opts + a, s, b + 1); return 0; }
PetscErrorCode PetscOptionsGetAll(void *o, char **copy) { (void)o; const char *s = g_opts ? g_opts : ""; *copy = malloc(strlen(s) + 1); strcpy(*copy, s); return 0; }
static int is_value(const char *t) { return t[0] != '-' || (t[1] >= '0' && t[1] <= '9') || t[1] == '.'; }
static int find_opt(const char *pre, const char *name, char *val, size_t len)
{
  char key[256]; snprintf(key, sizeof(key), "-%s%s", pre ? pre : "", name[0] == '-' ? name + 1 : name);
  if (!g_opts) return 0;
  char *dup = strdup(g_opts), *save = NULL; int found = 0;
  for (char *t = strtok_r(dup, " \t\n", &save); t; t = strtok_r(NULL, " \t\n", &save)) {
    if (!strcmp(t, key)) { found = 1; if (val) val[0] = 0; char *n = strtok_r(NULL, " \t\n", &save); if (n && is_value(n) && val) snprintf(val, len, "%s", n); if (!n) break; if (!is_value(n) && !strcmp(n, key)) continue; }
  }
  free(dup); return found;
}
PetscErrorCode PetscOptionsGetBool(void *o, const char *pre, const char *name, PetscBool *v, PetscBool *set)
{ (void)o; char val[64] = ""; int f = find_opt(pre, name, val, sizeof(val)); if (set) *set = f ? PETSC_TRUE : PETSC_FALSE; if (f) *v = (!val[0] || !strcmp(val, "1") || !strcmp(val, "true") || !strcmp(val, "yes")) ? PETSC_TRUE : PETSC_FALSE; return 0; }
PetscErrorCode PetscOptionsGetString(void *o, const char *pre, const char *name, char *buf, size_t len, PetscBool *set)
{ (void)o; int f = find_opt(pre, name, buf, len); if (set) *set = f ? PETSC_TRUE : PETSC_FALSE; return 0; }
PetscErrorCode PetscOptionsBool_Mock(PetscOptionItems *o, const char *name, PetscBool cur, PetscBool *v)
{ *v = cur; return PetscOptionsGetBool(NULL, o ? o->prefix : NULL, name, v, NULL); }

/* ---- PC: the driver side of the callback table (what KSPSetUp / KSPSolve do with a PC) */
PetscErrorCode PCCreate(MPI_Comm c, PC *pc) { (void)c; *pc = calloc(1, sizeof(**pc)); (*pc)->ops = &(*pc)->ops_storage; return 0; }
PetscErrorCode PCSetOptionsPrefix(PC pc, const char *p) { snprintf(pc->prefix, 64, "%s", p ? p : ""); return 0; }
PetscErrorCode PCGetOptionsPrefix(PC pc, const char **p) { *p = pc->prefix; return 0; }
PetscErrorCode PCSetType(PC pc, PCType t)
{
  if (pc->ops->destroy) { PetscErrorCode e = pc->ops->destroy(pc); if (e) return e; memset(pc->ops, 0, sizeof(*pc->ops)); pc->data = NULL; }
  for (int i = 0; i < g_npc; ++i) if (!strcmp(g_pc[i].name, t)) { PetscErrorCode e = g_pc[i].create(pc); if (!e) snprintf(pc->type, 64, "%s", t); pc->setupcalled = 0; return e; }
  return PetscMockError(PETSC_ERR_ARG_WRONG, "Unable to find requested PC type %s", t);
}
PetscErrorCode PCSetOperators(PC pc, Mat A, Mat P) { pc->mat = A; pc->pmat = P; pc->setupcalled = 0; return 0; }
PetscErrorCode PCGetOperators(PC pc, Mat *A, Mat *P) { if (A) *A = pc->mat; if (P) *P = pc->pmat; return 0; }
PetscErrorCode PCSetFromOptions(PC pc)
{
  char t[64]; PetscBool set; PetscErrorCode e = PetscOptionsGetString(NULL, pc->prefix, "-pc_type", t, sizeof(t), &set); if (e) return e;
  if (set && strcmp(t, pc->type)) { e = PCSetType(pc, t); if (e) return e; }
  if (pc->ops->setfromoptions) { PetscOptionItems o = { pc->prefix }; return pc->ops->setfromoptions(&o, pc); }
  return 0;
}
PetscErrorCode PCSetUp(PC pc)
{
  if (!pc->mat) return PetscMockError(PETSC_ERR_ORDER, "Matrix must be set first");
  if (pc->setupcalled) return 0;
  if (pc->ops->setup) { PetscErrorCode e = pc->ops->setup(pc); if (e) return e; }
  pc->setupcalled = 1; return 0;
}
PetscErrorCode PCApply(PC pc, Vec x, Vec y)
{
  if (x == y) return PetscMockError(PETSC_ERR_ARG_WRONG, "x and y must be different vectors");
  PetscErrorCode e = PCSetUp(pc); if (e) return e;
  if (!pc->ops->apply) return PetscMockError(PETSC_ERR_SUP, "PC does not have apply");
  return pc->ops->apply(pc, x, y);
}
PetscErrorCode PCView(PC pc, PetscViewer v)
{
  PetscErrorCode e = PetscViewerASCIIPrintf(v, "PC Object: (%s) 1 MPI processes\n  type: %s\n", pc->prefix, pc->type); if (e) return e;
  return pc->ops->view ? pc->ops->view(pc, v) : 0;
}
PetscErrorCode PCReset(PC pc) { PetscErrorCode e = 0; if (pc->ops->reset) e = pc->ops->reset(pc); pc->setupcalled = 0; return e; }
PetscErrorCode PCDestroy(PC *pc)
{ PetscErrorCode e = 0; if (*pc) { e = PCReset(*pc); if (!e && (*pc)->ops->destroy) e = (*pc)->ops->destroy(*pc); free(*pc); *pc = NULL; } return e; }

/* ---- viewer to a string */
PetscErrorCode PetscViewerStringOpen_Mock(PetscViewer *v) { *v = calloc(1, sizeof(**v)); (*v)->cap = 1 << 16; (*v)->buf = calloc(1, (*v)->cap); return 0; }
PetscErrorCode PetscViewerASCIIPrintf(PetscViewer v, const char *fmt, ...)
{ va_list ap; va_start(ap, fmt); int k = vsnprintf(v->buf + v->len, v->cap - v->len, fmt, ap); va_end(ap); if (k > 0) { v->len += (size_t)k; if (v->len >= v->cap) v->len = v->cap - 1; } return 0; }
PetscErrorCode PetscViewerDestroy(PetscViewer *v) { if (*v) { free((*v)->buf); free(*v); *v = NULL; } return 0; }
