{-# LANGUAGE ForeignFunctionInterface #-}
-- | Raw bindings to the shim surface of libjalla_b200.so (include/jalla_b200.h).
--   UNCOMPILED in the build image (no GHC there): kept mechanical so that it is a transcription
--   of the header.  The cblas_* / LAPACKE_* symbols need no new bindings — Jalla's existing
--   Numeric.Jalla.Foreign.BLAS / LAPACKE import them by name and resolve them in this library once
--   `extra-libraries` names it (see INTEGRATION.md).
module Numeric.Jalla.Foreign.B200
  ( jbInit, jbMalloc, jbFreePtr, jbMemcpyH2D, jbMemcpyD2H, jbSync, jbLastError
  , jbDlacpy, jbSlacpy, jbClacpy, jbZlacpy, jbLacpy, jbLaset, jbSetDiag, jbMultDiag, isComplex
  , jbDgesvBatched, jbDgetrfBatched, jbDgetrsBatched, jbDsolveBatched
  , memDevice, memManaged, memPinned
  ) where

import Foreign
import Foreign.C.Types
import Data.Complex (Complex)

memDevice, memManaged, memPinned :: CInt
memDevice = 0; memManaged = 1; memPinned = 2

foreign import ccall unsafe "jb_init"        jbInit      :: CInt -> IO CInt
foreign import ccall unsafe "jb_malloc"      jbMalloc    :: CSize -> CInt -> IO (Ptr a)
-- | Finalizer for ForeignPtr: thread-safe, context-agnostic (GHC runs finalizers on any thread).
foreign import ccall unsafe "&jb_free"       jbFreePtr   :: FunPtr (Ptr a -> IO ())
foreign import ccall unsafe "jb_memcpy_h2d"  jbMemcpyH2D :: Ptr a -> Ptr a -> CSize -> IO CInt
foreign import ccall unsafe "jb_memcpy_d2h"  jbMemcpyD2H :: Ptr a -> Ptr a -> CSize -> IO CInt
foreign import ccall unsafe "jb_sync"        jbSync      :: IO CInt
foreign import ccall unsafe "jb_last_error"  jbLastError :: IO CInt

-- trans (111 copy / 112 transpose / 113 conj-transpose), m, n, a, lda, b, ldb
foreign import ccall unsafe "jb_slacpy" jbSlacpy :: CInt -> CInt -> CInt -> Ptr CFloat  -> CInt -> Ptr CFloat  -> CInt -> IO CInt
foreign import ccall unsafe "jb_dlacpy" jbDlacpy :: CInt -> CInt -> CInt -> Ptr CDouble -> CInt -> Ptr CDouble -> CInt -> IO CInt
foreign import ccall unsafe "jb_clacpy" jbClacpy :: CInt -> CInt -> CInt -> Ptr a -> CInt -> Ptr a -> CInt -> IO CInt
foreign import ccall unsafe "jb_zlacpy" jbZlacpy :: CInt -> CInt -> CInt -> Ptr a -> CInt -> Ptr a -> CInt -> IO CInt

-- element-size dispatch for the polymorphic wrappers of Numeric.Jalla.B200 (4 = s, 8 real = d, 8 complex = c, 16 = z)
class IsComplex e where isComplex :: e -> Bool
instance IsComplex CFloat where isComplex _ = False
instance IsComplex CDouble where isComplex _ = False
instance IsComplex (Complex a) where isComplex _ = True

jbLacpy :: Int -> Bool -> CInt -> CInt -> CInt -> Ptr () -> CInt -> Ptr () -> CInt -> IO CInt
jbLacpy 4 False = \t m n a lda b ldb -> jbSlacpy t m n (castPtr a) lda (castPtr b) ldb
jbLacpy 8 False = \t m n a lda b ldb -> jbDlacpy t m n (castPtr a) lda (castPtr b) ldb
jbLacpy 8 True  = jbClacpy
jbLacpy _ _     = jbZlacpy

-- off-diagonal and diagonal values travel by pointer so that one binding serves all four element types
foreign import ccall unsafe "jb_claset" jbClaset :: CInt -> CInt -> Ptr () -> Ptr () -> Ptr () -> CInt -> IO CInt
foreign import ccall unsafe "jb_zlaset" jbZlaset :: CInt -> CInt -> Ptr () -> Ptr () -> Ptr () -> CInt -> IO CInt
foreign import ccall unsafe "jb_slaset" jbSlaset :: CInt -> CInt -> CFloat -> CFloat -> Ptr CFloat -> CInt -> IO CInt
foreign import ccall unsafe "jb_dlaset" jbDlaset :: CInt -> CInt -> CDouble -> CDouble -> Ptr CDouble -> CInt -> IO CInt
jbLaset :: Int -> Bool -> CInt -> CInt -> Ptr () -> Ptr () -> Ptr () -> CInt -> IO CInt
jbLaset 4 False m n po pd a lda = do { o <- peek (castPtr po); d <- peek (castPtr pd); jbSlaset m n o d (castPtr a) lda }
jbLaset 8 False m n po pd a lda = do { o <- peek (castPtr po); d <- peek (castPtr pd); jbDlaset m n o d (castPtr a) lda }
jbLaset 8 True  m n po pd a lda = jbClaset m n po pd a lda
jbLaset _ _     m n po pd a lda = jbZlaset m n po pd a lda

-- m, n, diagonal index, values (host or device), count, a, lda
foreign import ccall unsafe "jb_sset_diag" jbSsetDiag :: CInt -> CInt -> CInt -> Ptr () -> CInt -> Ptr () -> CInt -> IO CInt
foreign import ccall unsafe "jb_dset_diag" jbDsetDiag :: CInt -> CInt -> CInt -> Ptr () -> CInt -> Ptr () -> CInt -> IO CInt
foreign import ccall unsafe "jb_cset_diag" jbCsetDiag :: CInt -> CInt -> CInt -> Ptr () -> CInt -> Ptr () -> CInt -> IO CInt
foreign import ccall unsafe "jb_zset_diag" jbZsetDiag :: CInt -> CInt -> CInt -> Ptr () -> CInt -> Ptr () -> CInt -> IO CInt
jbSetDiag :: Int -> Bool -> CInt -> CInt -> CInt -> Ptr () -> CInt -> Ptr () -> CInt -> IO CInt
jbSetDiag 4 False = jbSsetDiag
jbSetDiag 8 False = jbDsetDiag
jbSetDiag 8 True  = jbCsetDiag
jbSetDiag _ _     = jbZsetDiag

-- trans, r, c, a, lda, d (host or device, c entries), out, ldo:  out = op(a) * diag(d)
foreign import ccall unsafe "jb_smult_diag" jbSmultDiag :: CInt -> CInt -> CInt -> Ptr () -> CInt -> Ptr () -> Ptr () -> CInt -> IO CInt
foreign import ccall unsafe "jb_dmult_diag" jbDmultDiag :: CInt -> CInt -> CInt -> Ptr () -> CInt -> Ptr () -> Ptr () -> CInt -> IO CInt
foreign import ccall unsafe "jb_cmult_diag" jbCmultDiag :: CInt -> CInt -> CInt -> Ptr () -> CInt -> Ptr () -> Ptr () -> CInt -> IO CInt
foreign import ccall unsafe "jb_zmult_diag" jbZmultDiag :: CInt -> CInt -> CInt -> Ptr () -> CInt -> Ptr () -> Ptr () -> CInt -> IO CInt
jbMultDiag :: Int -> Bool -> CInt -> CInt -> CInt -> Ptr () -> CInt -> Ptr () -> Ptr () -> CInt -> IO CInt
jbMultDiag 4 False = jbSmultDiag
jbMultDiag 8 False = jbDmultDiag
jbMultDiag 8 True  = jbCmultDiag
jbMultDiag _ _     = jbZmultDiag

-- batched 32x32-class systems (n <= 32): strided batch, device pointers
foreign import ccall unsafe "jb_dgetrf_batched" jbDgetrfBatched
  :: CInt -> Ptr CDouble -> CInt -> CLLong -> Ptr CInt -> Ptr CInt -> CInt -> IO CInt
foreign import ccall unsafe "jb_dgetrs_batched" jbDgetrsBatched
  :: CInt -> CInt -> Ptr CDouble -> CInt -> CLLong -> Ptr CInt -> Ptr CDouble -> CInt -> CLLong -> CInt -> IO CInt
foreign import ccall unsafe "jb_dgesv_batched" jbDgesvBatched
  :: CInt -> CInt -> Ptr CDouble -> CInt -> CLLong -> Ptr CInt -> Ptr CDouble -> CInt -> CLLong -> Ptr CInt -> CInt -> IO CInt
foreign import ccall unsafe "jb_dsolve_batched" jbDsolveBatched
  :: CInt -> CInt -> Ptr CDouble -> CInt -> CLLong -> Ptr CDouble -> CInt -> CLLong -> Ptr CDouble -> CInt -> CLLong -> Ptr CInt -> CInt -> IO CInt
