/* petsc_mock.h -- a minimal stand-in for the slice of PETSc that shim/pcexsaddleb200.c touches (TEST INFRASTRUCTURE).
 *
 * PETSc is absent from the build image, so the shim cannot be compiled against the real headers here.  This mock declares
 * the same names with the same argument meaning (PC / Mat / Vec objects, pc->ops slots of petsc/private/pcimpl.h, MATSHELL,
 * PCRegister / MatRegister, the options database, ASCII viewers, CHKERRQ / SETERRQ) and petsc_mock.c implements just enough
 * behaviour for a driver to walk a PC through create -> setfromoptions -> setup -> apply -> view -> reset -> destroy and a
 * Mat through create -> settype -> mult -> destroy, the way KSPSolve would.  With a real PETSc the shim is compiled WITHOUT
 * -DXSB_MOCK_PETSC and includes <petsc/private/pcimpl.h> instead (reference: pcildl.c:21).
 */
#ifndef PETSC_MOCK_H_
#define PETSC_MOCK_H_
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int PetscErrorCode;
typedef int PetscInt;
typedef double PetscScalar;
typedef double PetscReal;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscBool;
typedef int MPI_Comm;
#define PETSC_COMM_SELF 1
#define PETSC_COMM_WORLD 2
#define PETSC_EXTERN extern
#define PETSC_ERR_SUP 56
#define PETSC_ERR_USER 83
#define PETSC_ERR_LIB 76
#define PETSC_ERR_ORDER 58
#define PETSC_ERR_ARG_WRONG 62
#define PetscFunctionBegin
#define PetscFunctionReturn(x) return (x)
#define CHKERRQ(ierr) do { if (ierr) return (ierr); } while (0)
PetscErrorCode PetscMockError(int code, const char *fmt, ...);
const char *PetscMockLastError(void);
#define SETERRQ(comm, code, msg) return PetscMockError(code, "%s", msg)
#define SETERRQ1(comm, code, fmt, a) return PetscMockError(code, fmt, a)
#define SETERRQ2(comm, code, fmt, a, b) return PetscMockError(code, fmt, a, b)

typedef struct _p_Vec *Vec;
typedef struct _p_Mat *Mat;
typedef struct _p_PC *PC;
typedef struct _p_PetscViewer *PetscViewer;
typedef struct _p_PetscOptionItems PetscOptionItems;
typedef void *PetscObject;
typedef const char *MatType;
typedef const char *PCType;
typedef enum { MATOP_MULT = 3, MATOP_GET_DIAGONAL = 17, MATOP_DESTROY = 60 } MatOperation;
#define MATSHELL "shell"
#define PETSCVIEWERASCII "ascii"

struct _p_Vec { PetscInt n; PetscScalar *a; int rd, wr; };
struct _p_Mat { char type[64]; PetscInt m, n; void *shellctx; PetscErrorCode (*mult)(Mat, Vec, Vec); PetscErrorCode (*getdiagonal)(Mat, Vec); PetscErrorCode (*destroy)(Mat); };
struct _p_PetscViewer { char *buf; size_t len, cap; };
struct _p_PetscOptionItems { const char *prefix; };

/* the slots of struct _PCOps the reference fills (pcildl.c:469-478) */
struct _PCOps {
  PetscErrorCode (*setup)(PC);
  PetscErrorCode (*apply)(PC, Vec, Vec);
  PetscErrorCode (*applyrichardson)(PC, Vec, Vec, Vec, PetscReal, PetscReal, PetscReal, PetscInt, PetscBool, PetscInt *, int *);
  PetscErrorCode (*applytranspose)(PC, Vec, Vec);
  PetscErrorCode (*applysymmetricleft)(PC, Vec, Vec);
  PetscErrorCode (*applysymmetricright)(PC, Vec, Vec);
  PetscErrorCode (*setfromoptions)(PetscOptionItems *, PC);
  PetscErrorCode (*reset)(PC);
  PetscErrorCode (*destroy)(PC);
  PetscErrorCode (*view)(PC, PetscViewer);
};
struct _p_PC { struct _PCOps ops_storage, *ops; void *data; Mat mat, pmat; char prefix[64]; char type[64]; int setupcalled; };

/* memory */
#define PetscNewLog(obj, pp) ((*(pp) = calloc(1, sizeof(**(pp)))) ? 0 : 55)
#define PetscFree(p) (free(p), (p) = NULL, 0)
/* Vec */
PetscErrorCode VecCreateSeq(MPI_Comm, PetscInt n, Vec *);
PetscErrorCode VecDestroy(Vec *);
PetscErrorCode VecGetSize(Vec, PetscInt *);
PetscErrorCode VecGetArrayRead(Vec, const PetscScalar **);
PetscErrorCode VecRestoreArrayRead(Vec, const PetscScalar **);
PetscErrorCode VecGetArray(Vec, PetscScalar **);
PetscErrorCode VecRestoreArray(Vec, PetscScalar **);
/* Mat */
PetscErrorCode MatRegister(const char *, PetscErrorCode (*)(Mat));
PetscErrorCode MatCreate(MPI_Comm, Mat *);
PetscErrorCode MatSetSizes(Mat, PetscInt, PetscInt, PetscInt, PetscInt);
PetscErrorCode MatGetSize(Mat, PetscInt *, PetscInt *);
PetscErrorCode MatSetType(Mat, MatType);
PetscErrorCode MatShellSetContext(Mat, void *);
PetscErrorCode MatShellGetContext(Mat, void *);
PetscErrorCode MatShellSetOperation(Mat, MatOperation, void (*)(void));
PetscErrorCode MatMult(Mat, Vec, Vec);
PetscErrorCode MatGetDiagonal(Mat, Vec);
PetscErrorCode MatDestroy(Mat *);
PetscErrorCode PetscObjectTypeCompare(PetscObject, const char *, PetscBool *);
/* PC */
PetscErrorCode PCRegister(const char *, PetscErrorCode (*)(PC));
PetscErrorCode PCCreate(MPI_Comm, PC *);
PetscErrorCode PCSetOptionsPrefix(PC, const char *);
PetscErrorCode PCGetOptionsPrefix(PC, const char **);
PetscErrorCode PCSetType(PC, PCType);
PetscErrorCode PCSetOperators(PC, Mat, Mat);
PetscErrorCode PCGetOperators(PC, Mat *, Mat *);
PetscErrorCode PCSetFromOptions(PC);
PetscErrorCode PCSetUp(PC);
PetscErrorCode PCApply(PC, Vec, Vec);
PetscErrorCode PCView(PC, PetscViewer);
PetscErrorCode PCReset(PC);
PetscErrorCode PCDestroy(PC *);
/* options database */
PetscErrorCode PetscOptionsClear(void *);
PetscErrorCode PetscOptionsInsertString(void *, const char *);
PetscErrorCode PetscOptionsGetAll(void *, char **);   /* caller frees with PetscFree */
PetscErrorCode PetscOptionsGetBool(void *, const char *pre, const char *name, PetscBool *v, PetscBool *set);
PetscErrorCode PetscOptionsGetString(void *, const char *pre, const char *name, char *buf, size_t len, PetscBool *set);
#define PetscOptionsHead(o, s) 0
#define PetscOptionsTail() 0
PetscErrorCode PetscOptionsBool_Mock(PetscOptionItems *, const char *name, PetscBool cur, PetscBool *v);
#define PetscOptionsBool(name, text, man, cur, v, set) PetscOptionsBool_Mock(PetscOptionsObject, name, cur, v)
/* viewer */
PetscErrorCode PetscViewerStringOpen_Mock(PetscViewer *);
PetscErrorCode PetscViewerASCIIPrintf(PetscViewer, const char *fmt, ...);
PetscErrorCode PetscViewerDestroy(PetscViewer *);
#endif
