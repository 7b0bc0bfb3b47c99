/* pcexsaddleb200.c -- the reference-side binding of libexsaddle_b200.so: a PETSc MatType and a PETSc PCType.
 *
 * This is the file a maintainer adds to psanan/exsaddle (next to pcildl.c, which it follows slot for slot:
 * pcildl.c:460-485 creator, :289-322 setup, :326-372 apply, :376-394 reset, :398-407 destroy, :411-423 setfromoptions,
 * :427-456 view).  It is real C against PETSc's public + private PC API; in this repository it is compiled with
 * -DXSB_MOCK_PETSC against tests/mock_petsc (PETSc is absent from the image) and driven by tests/test_shim.py.
 *
 *   MATEXSADDLEB200  a MATSHELL whose context OWNS the xsb_ctx handle.  Created at the seam where the reference
 *                    chooses its matrix type (DMCreateMatrix_SaddleAIJ, femixedspace.c:628-642: the -viennaclhack2 /
 *                    -sbaijhack MatSetType calls).  The element loops of MatAssemble_Saddle are replaced by
 *                    MatAssemble_ExSaddleB200 -> xsb_assemble (GPU assembly from the same options); MatMult ->
 *                    xsb_mat_mult; MatGetDiagonal -> xsb_mat_get_diagonal.
 *   PCEXSADDLEB200   the ABF preconditioner (fieldsplit Schur-upper + GCR/GMG + ILU(0), abf.opts) as ONE PC.  It does
 *                    not create a handle: at PCSetUp it looks the Mat's handle up from pc->pmat (PCGetOperators +
 *                    MatShellGetContext), so Mat and PC share one xsb_ctx and one assembled operator.
 * Options: the whole PETSc options database is forwarded verbatim (xsb_set_options), so -mx / -model / -eta1 /
 * -saddle_fieldsplit_* / -options_file abf.opts drive the library exactly as they drive the reference.
 */
#ifdef XSB_MOCK_PETSC
#include "petsc_mock.h"
#else
#include <petsc/private/pcimpl.h>   /* pcildl.c:21 */
#include <petsc/private/matimpl.h>
#endif
#include "pcexsaddleb200.h"
#include "exsaddle_b200.h"

#ifndef NSD
#define NSD 3                        /* exSaddle.h:7-9: -DNSD=2|3, -DLAME select the executable */
#endif
#ifdef LAME
#define XSB_LAME 1
#else
#define XSB_LAME 0
#endif

#define XSB_MAT_MAGIC 0x58534232     /* "XSB2": tells our MATSHELL context from somebody else's */
typedef struct { int magic; xsb_ctx ctx; } Mat_ExSaddleB200;
typedef struct { xsb_ctx ctx; /* borrowed from pc->pmat */ PetscBool matrix_free; PetscBool setup; } PC_ExSaddleB200;

#define XSBCHK(h, call) do { int rc_ = (call); if (rc_) SETERRQ2(PETSC_COMM_SELF, PETSC_ERR_LIB, "exsaddle_b200 error %d: %s", rc_, xsb_last_error(h)); } while (0)

static PetscErrorCode ForwardOptions(xsb_ctx h)
{
  char *all = NULL; PetscErrorCode ierr;
  PetscFunctionBegin;
  ierr = PetscOptionsGetAll(NULL, &all);CHKERRQ(ierr);     /* "-mx 64 -model 6 -saddle_fieldsplit_u_pc_type mg ..." */
  if (all) { int rc = xsb_set_options(h, all); ierr = PetscFree(all);CHKERRQ(ierr); if (rc) SETERRQ2(PETSC_COMM_SELF, PETSC_ERR_LIB, "exsaddle_b200 error %d: %s", rc, xsb_last_error(h)); }
  PetscFunctionReturn(0);
}

/* ------------------------------------------------------------------------------------------- MatType */
static PetscErrorCode MatGetHandle(Mat A, xsb_ctx *h)
{
  Mat_ExSaddleB200 *m = NULL; PetscErrorCode ierr;
  PetscFunctionBegin;
  ierr = MatShellGetContext(A, &m);CHKERRQ(ierr);
  if (!m || m->magic != XSB_MAT_MAGIC) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_USER, "Only valid for Mat type " MATEXSADDLEB200);   /* as pcildl.c:318 */
  *h = m->ctx;
  PetscFunctionReturn(0);
}

PetscErrorCode MatAssemble_ExSaddleB200(Mat A)
{
  xsb_ctx h; PetscErrorCode ierr; int assembled = 0; int64_t sz[8]; PetscInt M, N;
  PetscFunctionBegin;
  ierr = MatGetHandle(A, &h);CHKERRQ(ierr);
  XSBCHK(h, xsb_get_state(h, &assembled, NULL));
  if (assembled) PetscFunctionReturn(0);
  ierr = ForwardOptions(h);CHKERRQ(ierr);
  XSBCHK(h, xsb_assemble(h));                              /* exSaddle.c:230-283 on the device */
  XSBCHK(h, xsb_get_sizes(h, sz));
  ierr = MatGetSize(A, &M, &N);CHKERRQ(ierr);
  if ((int64_t)M != sz[0] || (int64_t)N != sz[0]) SETERRQ2(PETSC_COMM_SELF, PETSC_ERR_USER, "Mat is %d x %d but the options describe a different system", (int)M, (int)N);
  PetscFunctionReturn(0);
}

static PetscErrorCode MatMult_ExSaddleB200(Mat A, Vec x, Vec y)
{
  xsb_ctx h; const PetscScalar *xa; PetscScalar *ya; PetscErrorCode ierr; int rc;
  PetscFunctionBegin;
  ierr = MatAssemble_ExSaddleB200(A);CHKERRQ(ierr);        /* lazily, if the caller did not */
  ierr = MatGetHandle(A, &h);CHKERRQ(ierr);
  ierr = VecGetArrayRead(x, &xa);CHKERRQ(ierr);            /* host arrays, as pcildl.c:335-336 */
  ierr = VecGetArray(y, &ya);CHKERRQ(ierr);
  rc = xsb_mat_mult(h, XSB_MAT_A, xa, ya);
  ierr = VecRestoreArrayRead(x, &xa);CHKERRQ(ierr);
  ierr = VecRestoreArray(y, &ya);CHKERRQ(ierr);
  if (rc) SETERRQ2(PETSC_COMM_SELF, PETSC_ERR_LIB, "exsaddle_b200 error %d: %s", rc, xsb_last_error(h));
  PetscFunctionReturn(0);
}
/* with a CUDA-enabled PETSc: VecCUDAGetArrayRead / VecCUDAGetArray + xsb_mat_mult_dev, no host copies */

static PetscErrorCode MatGetDiagonal_ExSaddleB200(Mat A, Vec d)
{
  xsb_ctx h; PetscScalar *da; PetscErrorCode ierr; int rc;
  PetscFunctionBegin;
  ierr = MatAssemble_ExSaddleB200(A);CHKERRQ(ierr);
  ierr = MatGetHandle(A, &h);CHKERRQ(ierr);
  ierr = VecGetArray(d, &da);CHKERRQ(ierr);
  rc = xsb_mat_get_diagonal(h, XSB_MAT_A, da);
  ierr = VecRestoreArray(d, &da);CHKERRQ(ierr);
  if (rc) SETERRQ2(PETSC_COMM_SELF, PETSC_ERR_LIB, "exsaddle_b200 error %d: %s", rc, xsb_last_error(h));
  PetscFunctionReturn(0);
}

PetscErrorCode MatGetRHS_ExSaddleB200(Mat A, Vec F)        /* F of exSaddle.c:263-281, Dirichlet lifting included */
{
  xsb_ctx h; PetscScalar *fa; PetscErrorCode ierr; int rc;
  PetscFunctionBegin;
  ierr = MatAssemble_ExSaddleB200(A);CHKERRQ(ierr);
  ierr = MatGetHandle(A, &h);CHKERRQ(ierr);
  ierr = VecGetArray(F, &fa);CHKERRQ(ierr);
  rc = xsb_vec_get_rhs(h, fa);
  ierr = VecRestoreArray(F, &fa);CHKERRQ(ierr);
  if (rc) SETERRQ2(PETSC_COMM_SELF, PETSC_ERR_LIB, "exsaddle_b200 error %d: %s", rc, xsb_last_error(h));
  PetscFunctionReturn(0);
}

static PetscErrorCode MatDestroy_ExSaddleB200(Mat A)
{
  Mat_ExSaddleB200 *m = NULL; PetscErrorCode ierr;
  PetscFunctionBegin;
  ierr = MatShellGetContext(A, &m);CHKERRQ(ierr);
  if (m && m->magic == XSB_MAT_MAGIC) { xsb_destroy(&m->ctx); m->magic = 0; ierr = PetscFree(m);CHKERRQ(ierr); ierr = MatShellSetContext(A, NULL);CHKERRQ(ierr); }
  PetscFunctionReturn(0);
}

PetscErrorCode MatCreate_ExSaddleB200(Mat A)
{
  Mat_ExSaddleB200 *m; PetscErrorCode ierr;
  PetscFunctionBegin;
  ierr = MatSetType(A, MATSHELL);CHKERRQ(ierr);
  ierr = PetscNewLog(A, &m);CHKERRQ(ierr);
  m->magic = XSB_MAT_MAGIC;
  if (xsb_create(&m->ctx, NSD, XSB_LAME, -1)) { ierr = PetscFree(m);CHKERRQ(ierr); SETERRQ(PETSC_COMM_SELF, PETSC_ERR_LIB, "xsb_create failed"); }
  ierr = MatShellSetContext(A, m);CHKERRQ(ierr);
  ierr = MatShellSetOperation(A, MATOP_MULT, (void (*)(void))MatMult_ExSaddleB200);CHKERRQ(ierr);
  ierr = MatShellSetOperation(A, MATOP_GET_DIAGONAL, (void (*)(void))MatGetDiagonal_ExSaddleB200);CHKERRQ(ierr);
  ierr = MatShellSetOperation(A, MATOP_DESTROY, (void (*)(void))MatDestroy_ExSaddleB200);CHKERRQ(ierr);
  PetscFunctionReturn(0);
}

/* ------------------------------------------------------------------------------------------- PCType */
static PetscErrorCode PCSetUp_ExSaddleB200(PC pc)           /* pcildl.c:289-322 */
{
  PC_ExSaddleB200 *s = (PC_ExSaddleB200 *)pc->data; Mat A, P; PetscErrorCode ierr;
  PetscFunctionBegin;
  ierr = PCGetOperators(pc, &A, &P);CHKERRQ(ierr);
  ierr = MatGetHandle(P ? P : A, &s->ctx);CHKERRQ(ierr);     /* the Mat's handle: one xsb_ctx for Mat and PC */
  ierr = MatAssemble_ExSaddleB200(P ? P : A);CHKERRQ(ierr);
  ierr = ForwardOptions(s->ctx);CHKERRQ(ierr);               /* -saddle_fieldsplit_* / abf.opts */
  XSBCHK(s->ctx, xsb_set_option(s->ctx, "-fs", NULL));       /* this PC IS the -fs tree of exSaddle.c:303-322 */
  if (s->matrix_free) XSBCHK(s->ctx, xsb_set_option(s->ctx, "-xsb_matrix_free", NULL));
  XSBCHK(s->ctx, xsb_ksp_setup(s->ctx));
  s->setup = PETSC_TRUE;
  PetscFunctionReturn(0);
}

static PetscErrorCode PCApply_ExSaddleB200(PC pc, Vec b, Vec x)   /* pcildl.c:326-372 */
{
  PC_ExSaddleB200 *s = (PC_ExSaddleB200 *)pc->data; const PetscScalar *ba; PetscScalar *xa; PetscErrorCode ierr; int rc;
  PetscFunctionBegin;
  if (!s->setup) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ORDER, "PCApply_ExSaddleB200 before PCSetUp");
  ierr = VecGetArrayRead(b, &ba);CHKERRQ(ierr);
  ierr = VecGetArray(x, &xa);CHKERRQ(ierr);
  rc = xsb_pc_apply(s->ctx, ba, xa);
  ierr = VecRestoreArrayRead(b, &ba);CHKERRQ(ierr);
  ierr = VecRestoreArray(x, &xa);CHKERRQ(ierr);
  if (rc) SETERRQ2(PETSC_COMM_SELF, PETSC_ERR_LIB, "exsaddle_b200 error %d: %s", rc, xsb_last_error(s->ctx));
  PetscFunctionReturn(0);
}

static PetscErrorCode PCReset_ExSaddleB200(PC pc)           /* pcildl.c:376-394: NULL-checked, idempotent */
{
  PC_ExSaddleB200 *s = (PC_ExSaddleB200 *)pc->data;
  PetscFunctionBegin;
  if (s && s->ctx && s->setup) xsb_ksp_reset(s->ctx);      /* solver state only: the operator belongs to the Mat */
  if (s) { s->setup = PETSC_FALSE; s->ctx = NULL; }
  PetscFunctionReturn(0);
}

static PetscErrorCode PCDestroy_ExSaddleB200(PC pc)         /* pcildl.c:398-407 */
{
  PetscErrorCode ierr;
  PetscFunctionBegin;
  ierr = PCReset_ExSaddleB200(pc);CHKERRQ(ierr);
  ierr = PetscFree(pc->data);CHKERRQ(ierr);
  PetscFunctionReturn(0);
}

static PetscErrorCode PCSetFromOptions_ExSaddleB200(PetscOptionItems *PetscOptionsObject, PC pc)   /* pcildl.c:411-423 */
{
  PC_ExSaddleB200 *s = (PC_ExSaddleB200 *)pc->data; PetscErrorCode ierr;
  PetscFunctionBegin;
  ierr = PetscOptionsHead(PetscOptionsObject, "ExSaddleB200 options");CHKERRQ(ierr);
  ierr = PetscOptionsBool("-pc_exsaddleb200_matrix_free", "Fine-level A00 products by the element kernel", NULL, s->matrix_free, &s->matrix_free, NULL);CHKERRQ(ierr);
  ierr = PetscOptionsTail();CHKERRQ(ierr);
  PetscFunctionReturn(0);
}

static PetscErrorCode PCView_ExSaddleB200(PC pc, PetscViewer viewer)   /* pcildl.c:427-456 */
{
  PC_ExSaddleB200 *s = (PC_ExSaddleB200 *)pc->data; PetscErrorCode ierr; PetscBool iascii;
  PetscFunctionBegin;
  ierr = PetscObjectTypeCompare((PetscObject)viewer, PETSCVIEWERASCII, &iascii);CHKERRQ(ierr);
  if (iascii) {
    ierr = PetscViewerASCIIPrintf(viewer, "  ExSaddleB200: matrix_free : %d\n", (int)s->matrix_free);CHKERRQ(ierr);
    if (s->setup) {
      static char buf[16384];
      if (!xsb_ksp_view(s->ctx, buf, (int)sizeof(buf))) { ierr = PetscViewerASCIIPrintf(viewer, "%s", buf);CHKERRQ(ierr); }
    } else { ierr = PetscViewerASCIIPrintf(viewer, "  ExSaddleB200: not yet set up\n");CHKERRQ(ierr); }
  }
  PetscFunctionReturn(0);
}

PetscErrorCode PCCreate_ExSaddleB200(PC pc)                 /* pcildl.c:460-485 */
{
  PC_ExSaddleB200 *s; PetscErrorCode ierr;
  PetscFunctionBegin;
  ierr     = PetscNewLog(pc, &s);CHKERRQ(ierr);
  pc->data = (void *)s;

  pc->ops->apply               = PCApply_ExSaddleB200;
  pc->ops->applytranspose      = 0;
  pc->ops->setup               = PCSetUp_ExSaddleB200;
  pc->ops->reset               = PCReset_ExSaddleB200;
  pc->ops->destroy             = PCDestroy_ExSaddleB200;
  pc->ops->setfromoptions      = PCSetFromOptions_ExSaddleB200;
  pc->ops->view                = PCView_ExSaddleB200;
  pc->ops->applyrichardson     = 0;
  pc->ops->applysymmetricleft  = 0;
  pc->ops->applysymmetricright = 0;

  s->ctx = NULL; s->matrix_free = PETSC_FALSE; s->setup = PETSC_FALSE;
  PetscFunctionReturn(0);
}
