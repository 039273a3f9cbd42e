{-# LANGUAGE ForeignFunctionInterface #-}
-- | Raw bindings to the shim surface of libjalla_b200.so (include/jalla_b200.h).
--   UNCOMPILED in the build image (no GHC there): kept mechanical so that it is a transcription
--   of the header.  The cblas_* / LAPACKE_* symbols need no new bindings — Jalla's existing
--   Numeric.Jalla.Foreign.BLAS / LAPACKE import them by name and resolve them in this library once
--   `extra-libraries` names it (see INTEGRATION.md).
module Numeric.Jalla.Foreign.B200
  ( jbInit, jbMalloc, jbFreePtr, jbMemcpyH2D, jbMemcpyD2H, jbSync, jbLastError
  , jbDlacpy, jbSlacpy, jbClacpy, jbZlacpy
  , jbDgesvBatched, jbDgetrfBatched, jbDgetrsBatched, jbDsolveBatched
  , memDevice, memManaged, memPinned
  ) where

import Foreign
import Foreign.C.Types

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

-- batched 32x32-class systems (n <= 32): strided batch, device pointers
foreign import ccall unsafe "jb_dgetrf_batched" jbDgetrfBatched
  :: CInt -> Ptr CDouble -> CInt -> CLLong -> Ptr CInt -> Ptr CInt -> CInt -> IO CInt
foreign import ccall unsafe "jb_dgetrs_batched" jbDgetrsBatched
  :: CInt -> CInt -> Ptr CDouble -> CInt -> CLLong -> Ptr CInt -> Ptr CDouble -> CInt -> CLLong -> CInt -> IO CInt
foreign import ccall unsafe "jb_dgesv_batched" jbDgesvBatched
  :: CInt -> CInt -> Ptr CDouble -> CInt -> CLLong -> Ptr CInt -> Ptr CDouble -> CInt -> CLLong -> Ptr CInt -> CInt -> IO CInt
foreign import ccall unsafe "jb_dsolve_batched" jbDsolveBatched
  :: CInt -> CInt -> Ptr CDouble -> CInt -> CLLong -> Ptr CDouble -> CInt -> CLLong -> Ptr CDouble -> CInt -> CLLong -> Ptr CInt -> CInt -> IO CInt
