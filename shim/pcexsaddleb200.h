/* pcexsaddleb200.h -- PETSc plug-in types backed by libexsaddle_b200.so (declarations; see pcexsaddleb200.c).
   Follows the reference's own plug-in header (pcildl.h:1-8): a type-name macro and the creator. */
#ifndef PCEXSADDLEB200_H_
#define PCEXSADDLEB200_H_
#ifdef XSB_MOCK_PETSC
#include "petsc_mock.h"
#else
#include <petscpc.h>
#endif

#define PCEXSADDLEB200  "exsaddleb200"
#define MATEXSADDLEB200 "exsaddleb200"

PETSC_EXTERN PetscErrorCode PCCreate_ExSaddleB200(PC);     /* PCRegister(PCEXSADDLEB200, PCCreate_ExSaddleB200), beside exSaddle.c:110-115 */
PETSC_EXTERN PetscErrorCode MatCreate_ExSaddleB200(Mat);   /* MatRegister(MATEXSADDLEB200, MatCreate_ExSaddleB200); selected in DMCreateMatrix_SaddleAIJ, femixedspace.c:628-642 */
/* what SaddleSolve_Q2Q1 calls instead of MatAssemble_Saddle_NULL / MatAssemble_Saddle / VecAssemble_F* / Impose... (exSaddle.c:267-281) */
PETSC_EXTERN PetscErrorCode MatAssemble_ExSaddleB200(Mat A);
PETSC_EXTERN PetscErrorCode MatGetRHS_ExSaddleB200(Mat A, Vec F);
#endif
