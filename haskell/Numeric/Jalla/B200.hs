{-# LANGUAGE MultiParamTypeClasses, FlexibleInstances, FlexibleContexts #-}
-- | GPU-resident matrix type for Jalla: a new instance of Jalla's own storage classes
--   (GMatrix / CMatrix, Numeric/Jalla/Matrix.hs:179-186,234-243) whose buffer lives in B200 HBM.
--   UNCOMPILED in the build image (no GHC there).  Everything arithmetic goes through the class
--   methods Jalla already has: gemm / gesv / getrf / getri dispatch on the ELEMENT type
--   (BlasOps e, LapackeOps e se), reach `cblas_?gemm` / `LAPACKE_?…` by name, and those symbols are
--   provided by libjalla_b200.so, which accepts device pointers.  So `##`, `##!`,
--   `solveLinearSystem` and `invert` work on GpuMatrix unchanged.
module Numeric.Jalla.B200 (GpuMatrix, toGpu, fromGpu) where

import Foreign
import Foreign.C.Types
import System.IO.Unsafe (unsafePerformIO)
import Numeric.Jalla.Matrix
import Numeric.Jalla.Types
import Numeric.Jalla.Foreign.BlasOps (BlasOps)
import Numeric.Jalla.Foreign.B200

-- | Same record as `data Matrix e` (Matrix.hs:329-332); only the allocator differs.
data GpuMatrix e = GpuMatrix { gpuPtr :: !(ForeignPtr e), gpuShape :: !Shape, gpuLDA :: !Index, gpuOrder :: !Order }

instance (BlasOps e) => GMatrix GpuMatrix e where
  shape = gpuShape
  -- element reads are device->host copies of one element (slow path; bulk traffic uses toGpu/fromGpu)
  m ! (i, j) = unsafePerformIO $ withForeignPtr (gpuPtr m) $ \p -> alloca $ \h -> do
    _ <- jbMemcpyD2H h (p `advancePtr` (i + j * gpuLDA m)) (fromIntegral (sizeOf (undefined :: e)))
    peek h

instance (BlasOps e) => CMatrix GpuMatrix e where
  -- replaces mallocForeignPtrArray (Matrix.hs:785-787): ColumnMajor, lda = rows, uninitialised
  matrixAlloc (r, c) = do
    p  <- jbMalloc (fromIntegral (max 1 (r * c) * sizeOf (undefined :: e))) memDevice
    fp <- newForeignPtr jbFreePtr p
    return (GpuMatrix fp (r, c) (max 1 r) ColumnMajor)
  withCMatrix m     = withForeignPtr (gpuPtr m)
  lda               = gpuLDA
  order             = gpuOrder
  matrixForeignPtr  = gpuPtr

-- matrixCopy on the device: one lacpy instead of r strided cblas_?copy calls (Matrix.hs:855-871);
-- wire it by giving GpuMatrix its own `matrixCopy`/`unsafeMatrixCopy` that call jb_?lacpy.

-- | Bulk host -> device (replaces per-element createMatrix/setElem loops, Matrix.hs:798-848).
toGpu :: (BlasOps e, Storable e) => Matrix e -> IO (GpuMatrix e)
toGpu m = do
  g <- matrixAlloc (shape m)
  let bytes = fromIntegral (colCount m * lda m * sizeOf (undefined `asTypeOf` (m ! (0, 0))))
  _ <- withCMatrix m $ \hp -> withCMatrix g $ \dp -> jbMemcpyH2D dp hp bytes
  return g

fromGpu :: (BlasOps e, Storable e) => GpuMatrix e -> IO (Matrix e)
fromGpu g = do
  m <- matrixAlloc (shape g)
  let bytes = fromIntegral (colCount g * lda g * sizeOf (undefined `asTypeOf` (m ! (0, 0))))
  _ <- withCMatrix g $ \dp -> withCMatrix m $ \hp -> jbMemcpyD2H hp dp bytes
  return m
