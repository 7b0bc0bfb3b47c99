/* drive_shim.c -- walks shim/pcexsaddleb200.c through the life cycle PETSc's KSP would (TEST INFRASTRUCTURE):
 *   PCRegister / MatRegister -> MatCreate + MatSetType -> PCCreate + PCSetOperators + PCSetFromOptions(-saddle_pc_type exsaddleb200)
 *   -> PCSetUp -> PCApply / MatMult (compared with the direct C-ABI calls on the same handle) -> PCView
 *   -> PCReset (twice) -> PCSetUp + PCApply again -> PCDestroy -> MatDestroy.
 * usage: drive_shim "<exSaddle options>" <rows>        Prints one "key: value" line per step; exit 0 when every step behaved. */
#include "petsc_mock.h"
#include "pcexsaddleb200.h"
#include "exsaddle_b200.h"
#include <math.h>

typedef struct { int magic; xsb_ctx ctx; } MatCtxView;   /* layout of the shell context (first two members) */

int main(int argc, char **argv)
{
  if (argc < 3) { fprintf(stderr, "usage: drive_shim \"options\" rows\n"); return 2; }
  const int n = atoi(argv[2]); PetscErrorCode ierr; Mat A = NULL; PC pc = NULL; Vec F = NULL, z = NULL, z2 = NULL, y = NULL; int bad = 0;
  ierr = PCRegister(PCEXSADDLEB200, PCCreate_ExSaddleB200); printf("PCRegister: %d\n", ierr); bad |= ierr;
  ierr = MatRegister(MATEXSADDLEB200, MatCreate_ExSaddleB200); printf("MatRegister: %d\n", ierr); bad |= ierr;
  PetscOptionsInsertString(NULL, argv[1]);
  ierr = MatCreate(PETSC_COMM_SELF, &A); ierr |= MatSetSizes(A, n, n, n, n); ierr |= MatSetType(A, MATEXSADDLEB200); printf("MatSetType: %d\n", ierr); bad |= ierr;
  ierr = PCCreate(PETSC_COMM_SELF, &pc); ierr |= PCSetOptionsPrefix(pc, "saddle_"); ierr |= PCSetOperators(pc, A, A); ierr |= PCSetFromOptions(pc);
  printf("PCSetFromOptions: %d type=%s\n", ierr, pc->type); bad |= ierr; bad |= strcmp(pc->type, PCEXSADDLEB200) != 0;
  { Vec a, b; VecCreateSeq(PETSC_COMM_SELF, n, &a); VecCreateSeq(PETSC_COMM_SELF, n, &b);   /* apply before set-up of a PC whose Mat cannot assemble, or plain success */
    PetscViewer v; PetscViewerStringOpen_Mock(&v); PCView(pc, v); printf("PCView(before setup): %s", strstr(v->buf, "not yet set up") ? "not yet set up\n" : "UNEXPECTED\n"); PetscViewerDestroy(&v);
    VecDestroy(&a); VecDestroy(&b); }
  ierr = PCSetUp(pc);
  printf("PCSetUp: %d%s%s\n", ierr, ierr ? " msg=" : "", ierr ? PetscMockLastError() : "");
  if (!ierr) {
    MatCtxView *m = NULL; MatShellGetContext(A, &m); xsb_ctx h = m->ctx;
    VecCreateSeq(PETSC_COMM_SELF, n, &F); VecCreateSeq(PETSC_COMM_SELF, n, &z); VecCreateSeq(PETSC_COMM_SELF, n, &z2); VecCreateSeq(PETSC_COMM_SELF, n, &y);
    ierr = MatGetRHS_ExSaddleB200(A, F); printf("MatGetRHS: %d\n", ierr); bad |= ierr;
    ierr = PCApply(pc, F, z); printf("PCApply: %d\n", ierr); bad |= ierr;
    int rc = xsb_pc_apply(h, F->a, z2->a); double d = 0, nz = 0; for (int i = 0; i < n; ++i) { d = fmax(d, fabs(z->a[i] - z2->a[i])); nz = fmax(nz, fabs(z->a[i])); }
    printf("PCApply vs xsb_pc_apply: rc=%d maxdiff=%.3e max=%.3e\n", rc, d, nz); bad |= rc || d != 0.0 || nz == 0.0;
    ierr = MatMult(A, z, y); rc = xsb_mat_mult(h, XSB_MAT_A, z->a, z2->a); d = 0; for (int i = 0; i < n; ++i) d = fmax(d, fabs(y->a[i] - z2->a[i]));
    printf("MatMult vs xsb_mat_mult: %d rc=%d maxdiff=%.3e\n", ierr, rc, d); bad |= ierr || rc || d != 0.0;
    ierr = MatGetDiagonal(A, y); printf("MatGetDiagonal: %d d[0]=%.6e\n", ierr, y->a[0]); bad |= ierr;
    { PetscViewer v; PetscViewerStringOpen_Mock(&v); ierr = PCView(pc, v); printf("PCView: %d fieldsplit=%d\n", ierr, strstr(v->buf, "FieldSplit with Schur preconditioner") != NULL); bad |= ierr || !strstr(v->buf, "FieldSplit"); PetscViewerDestroy(&v); }
    for (int i = 0; i < n; ++i) z2->a[i] = z->a[i];
    ierr = PCReset(pc); ierr |= PCReset(pc); printf("PCReset x2: %d\n", ierr); bad |= ierr;
    ierr = PCApply(pc, F, z);   /* PCApply sets the PC up again (setupcalled = 0 after reset) */
    d = 0; for (int i = 0; i < n; ++i) d = fmax(d, fabs(z->a[i] - z2->a[i]));
    printf("PCApply after reset: %d maxdiff=%.3e\n", ierr, d); bad |= ierr || d != 0.0;
  } else {
    ierr = PCReset(pc); ierr |= PCReset(pc); printf("PCReset x2: %d\n", ierr); bad |= ierr;
  }
  ierr = PCDestroy(&pc); printf("PCDestroy: %d\n", ierr); bad |= ierr;
  VecDestroy(&F); VecDestroy(&z); VecDestroy(&z2); VecDestroy(&y);
  ierr = MatDestroy(&A); printf("MatDestroy: %d\n", ierr); bad |= ierr;
  printf("lifecycle: %s\n", bad ? "FAILED" : "ok");
  return bad ? 1 : 0;
}
