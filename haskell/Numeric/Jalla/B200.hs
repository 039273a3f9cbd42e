{-# LANGUAGE MultiParamTypeClasses, FlexibleInstances, FlexibleContexts, TypeFamilies, ScopedTypeVariables #-}
-- | GPU-resident matrix type for Jalla: a new instance of Jalla's own classes whose buffer lives in B200 HBM.
--   It is written instance for instance like the reference's `Matrix` (Numeric/Jalla/Matrix.hs:329-357):
--
--   > instance … => GMatrix Matrix e          (:335-338)   ->  instance … => GMatrix GpuMatrix e
--   > instance BlasOps e => MatrixMatrix Matrix e (:340)   ->  instance BlasOps e => MatrixMatrix GpuMatrix e
--   > instance BlasOps e => Indexable (Matrix e) IndexPair e (:342-343)  ->  … (GpuMatrix e) …
--   > instance … => CMatrix Matrix e          (:346-353)   ->  instance … => CMatrix GpuMatrix e
--
--   UNCOMPILED in the build image (no GHC / cabal / c2hs there — SURVEY.md §8c), so it is kept a transcription of the
--   reference's own instance block with only the allocator and the element accessor changed.
--
--   What works unchanged on a GpuMatrix, and why: `##`, `##!` (class defaults, Matrix.hs:198-203) call matrixMult ->
--   unsafeMatrixMult -> BlasOps.gemm -> cblas_?gemm; solveLinearSystem / invert (Matrix.hs:513-540) call the top-level
--   matrixCopy — NOT a class method, so it cannot be overridden here — which is a loop of strided cblas_?copy calls
--   (Matrix.hs:855-871, CVector.hs:284-288) followed by LAPACKE_?gesv / ?getrf / ?getri.  All of these symbols resolve in
--   libjalla_b200.so, and every one of them — including cblas_?copy — accepts device pointers (include/jalla_b200.h,
--   "level 1, device-aware"; tests/cpp/solve_sequence.c replays exactly this sequence).  The per-row copy loop is r small
--   kernel launches; `gpuMatrixCopy` below is the one-launch form for callers that can choose.
--   What does NOT work on device memory: everything that peeks / pokes through `withCMatrix` on the host —
--   createMatrix / setElem / fill / setDiag (Matrix.hs:798-848,1013-1041), matrixMap, listMatrix.  Build on the host
--   and `toGpu`, use `gpuIdMatrix` / `gpuFill` / `gpuSetDiag` / `gpuMultDiag`, or allocate with `memManaged`.
module Numeric.Jalla.B200
  ( GpuMatrix, toGpu, fromGpu, gpuMatrixCopy, gpuIdMatrix, gpuFill, gpuSetDiag, gpuMultDiag ) where

import Foreign
import Foreign.C.Types
import System.IO.Unsafe (unsafePerformIO)
import Numeric.Jalla.Matrix
import Numeric.Jalla.CVector (Vector)
import Numeric.Jalla.Indexable
import Numeric.Jalla.Types
import Numeric.Jalla.Foreign.BlasOps (BlasOps)
import Numeric.Jalla.Foreign.B200

-- | Same record as `data Matrix e` (Matrix.hs:329-332); only the allocator differs.
data GpuMatrix e = GpuMatrix { gpuP :: !(ForeignPtr e), gpuShape :: !Shape, gpuLDA :: !Index, gpuOrder :: !Order }

instance (Num e, Field1 e, BlasOps e) => GMatrix GpuMatrix e where
  shape = gpuShape

instance BlasOps e => MatrixMatrix GpuMatrix e

-- | One element = one 8/16-byte device-to-host copy (the slow path, like the reference's matrixElem, Matrix.hs:893-899).
instance BlasOps e => Indexable (GpuMatrix e) IndexPair e where
  m ! (i, j)
    | not (i >= 0 && j >= 0 && i < fst (gpuShape m) && j < snd (gpuShape m)) = error "matrixElem out of bounds"
    | otherwise = unsafePerformIO $ withForeignPtr (gpuP m) $ \p -> alloca $ \h -> do
        let off = if gpuOrder m == RowMajor then i * gpuLDA m + j else j * gpuLDA m + i
        _ <- jbMemcpyD2H (castPtr h) (castPtr (p `advancePtr` off)) (fromIntegral (sizeOf (undefined :: e)))
        peek h

instance (Num e, Field1 e, BlasOps e) => CMatrix GpuMatrix e where
  type CMatrixVector GpuMatrix e  = Vector e                    -- host vectors, as for Matrix (Matrix.hs:347-348)
  type CMatrixVectorS GpuMatrix e = Vector (FieldScalar e)
  -- replaces matrixAlloc' = mallocForeignPtrArray (Matrix.hs:785-787): ColumnMajor, lda = rows, uninitialised
  matrixAlloc (r, c) = do
    p  <- jbMalloc (fromIntegral (max 1 (r * c) * sizeOf (undefined :: e))) memDevice
    fp <- newForeignPtr jbFreePtr p
    return (GpuMatrix fp (r, c) (max 1 r) ColumnMajor)
  withCMatrix m     = withForeignPtr (gpuP m)
  lda               = gpuLDA
  order             = gpuOrder
  matrixForeignPtr  = gpuP

-- | matrixCopy in one launch (jb_?lacpy) instead of r strided cblas_?copy calls.
gpuMatrixCopy :: forall e. (Num e, Field1 e, BlasOps e) => GpuMatrix e -> Transpose -> IO (GpuMatrix e)
gpuMatrixCopy a t = do
  ret <- matrixAlloc (shapeTrans t (shape a))
  let (r, c) = shape a
  _ <- withCMatrix a $ \pa -> withCMatrix ret $ \pr ->
         jbLacpy (sizeOf (undefined :: e)) (isComplex (undefined :: e)) (if t == NoTrans then 111 else 112)
                 (fromIntegral r) (fromIntegral c) (castPtr pa) (fromIntegral (lda a)) (castPtr pr) (fromIntegral (lda ret))
  return ret

-- | idMatrix / fill / setDiag on the device (Matrix.hs:525-526,1023-1036): jb_?laset, jb_?set_diag.
gpuFill :: forall e. (Num e, Field1 e, BlasOps e) => Shape -> e -> IO (GpuMatrix e)
gpuFill s v = do
  m <- matrixAlloc s
  _ <- withCMatrix m $ \p -> with v $ \pv ->
         jbLaset (sizeOf (undefined :: e)) (isComplex (undefined :: e)) (fromIntegral (fst s)) (fromIntegral (snd s)) (castPtr pv) (castPtr pv) (castPtr p) (fromIntegral (lda m))
  return m

gpuIdMatrix :: forall e. (Num e, Field1 e, BlasOps e) => Index -> IO (GpuMatrix e)
gpuIdMatrix n = do
  m <- matrixAlloc (n, n)
  _ <- withCMatrix m $ \p -> with (0 :: e) $ \p0 -> with (1 :: e) $ \p1 ->
         jbLaset (sizeOf (undefined :: e)) (isComplex (undefined :: e)) (fromIntegral n) (fromIntegral n) (castPtr p0) (castPtr p1) (castPtr p) (fromIntegral (lda m))
  return m

gpuSetDiag :: forall e. (Num e, Field1 e, BlasOps e) => GpuMatrix e -> Index -> [e] -> IO ()
gpuSetDiag m d vs = withArrayLen vs $ \k pv -> withCMatrix m $ \p -> do
  let (r, c) = shape m
  _ <- jbSetDiag (sizeOf (undefined :: e)) (isComplex (undefined :: e)) (fromIntegral r) (fromIntegral c) (fromIntegral d) (castPtr pv) (fromIntegral k) (castPtr p) (fromIntegral (lda m))
  return ()

-- | matrixMultDiag (Matrix.hs:660-665) in one launch: op(A) * diag(d).
gpuMultDiag :: forall e. (Num e, Field1 e, BlasOps e) => (GpuMatrix e, Transpose) -> [e] -> IO (GpuMatrix e)
gpuMultDiag (a, t) d = do
  let (r, c) = shapeTrans t (shape a)
  out <- matrixAlloc (r, c)
  _ <- withArrayLen (take c d) $ \_ pd -> withCMatrix a $ \pa -> withCMatrix out $ \po ->
         jbMultDiag (sizeOf (undefined :: e)) (isComplex (undefined :: e)) (if t == NoTrans then 111 else 112) (fromIntegral r) (fromIntegral c)
                    (castPtr pa) (fromIntegral (lda a)) (castPtr pd) (castPtr po) (fromIntegral (lda out))
  return out

-- | Bulk host <-> device (replaces per-element createMatrix / matrixList traffic, Matrix.hs:798-848,893-918).
toGpu :: forall e. (Num e, Field1 e, BlasOps e) => Matrix e -> IO (GpuMatrix e)
toGpu m = do
  g <- matrixAlloc (shape m)
  let bytes = fromIntegral (colCount m * lda m * sizeOf (undefined :: e))
  _ <- withCMatrix m $ \hp -> withCMatrix g $ \dp -> jbMemcpyH2D (castPtr dp) (castPtr hp) bytes
  return g

fromGpu :: forall e. (Num e, Field1 e, BlasOps e) => GpuMatrix e -> IO (Matrix e)
fromGpu g = do
  m <- matrixAlloc (shape g)
  let bytes = fromIntegral (colCount g * lda g * sizeOf (undefined :: e))
  _ <- withCMatrix g $ \dp -> withCMatrix m $ \hp -> jbMemcpyD2H (castPtr hp) (castPtr dp) bytes
  return m
