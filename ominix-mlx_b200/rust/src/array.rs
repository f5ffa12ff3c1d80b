//! A strided view of device memory + the stream it is used on.
use crate::error::{guard, Exception, Result};
use crate::ffi;
use std::os::raw::c_void;
use std::sync::Arc;

#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Dtype {
    Bool,
    Int32,
    Float16,
    Float32,
    Bfloat16,
}

impl Dtype {
    pub fn size(self) -> usize {
        match self {
            Dtype::Bool => 1,
            Dtype::Float16 | Dtype::Bfloat16 => 2,
            Dtype::Int32 | Dtype::Float32 => 4,
        }
    }
    pub(crate) fn code(self) -> i32 {
        match self {
            Dtype::Bool => 0,
            Dtype::Int32 => 7,
            Dtype::Float16 => 9,
            Dtype::Float32 => 10,
            Dtype::Bfloat16 => 12,
        }
    }
    pub(crate) fn from_code(c: i32) -> Self {
        match c {
            0 => Dtype::Bool,
            7 => Dtype::Int32,
            9 => Dtype::Float16,
            10 => Dtype::Float32,
            _ => Dtype::Bfloat16,
        }
    }
}

/// A cudaStream_t.  `Stream::default()` is the legacy default stream, the analogue of
/// `Stream::task_local_or_default()` that `#[default_device]` injects in mlx-rs.
#[derive(Clone, Copy, Debug)]
pub struct Stream(pub *mut c_void);
impl Default for Stream {
    fn default() -> Self {
        Stream(std::ptr::null_mut())
    }
}
unsafe impl Send for Stream {}

/// Device allocation owned by an `Array` (freed on drop) or borrowed from a cache.
pub(crate) enum Storage {
    Owned { ptr: *mut c_void, free: unsafe fn(*mut c_void) },
    Borrowed(#[allow(dead_code)] Arc<dyn std::any::Any + Send + Sync>),
    External,
}
impl Drop for Storage {
    fn drop(&mut self) {
        if let Storage::Owned { ptr, free } = self {
            unsafe { free(*ptr) }
        }
    }
}

/// Strided device array, `Send` but not `Sync` like `mlx_rs::Array` (mlx-rs/src/array/mod.rs:74).
#[derive(Clone)]
pub struct Array {
    pub(crate) desc: ffi::omx_array,
    pub(crate) _storage: Arc<Storage>,
}
unsafe impl Send for Array {}

extern "C" {
    fn cudaMalloc(ptr: *mut *mut c_void, size: usize) -> i32;
    fn cudaFree(ptr: *mut c_void) -> i32;
}
unsafe fn cuda_free(p: *mut c_void) {
    cudaFree(p);
}

impl Array {
    /// Wrap memory owned elsewhere (e.g. a projection output of the host framework).
    ///
    /// # Safety
    /// `data` must stay valid for the lifetime of the returned view.
    pub unsafe fn from_raw(data: *mut c_void, dtype: Dtype, shape: &[i64], strides: &[i64]) -> Result<Self> {
        if shape.len() != strides.len() || shape.len() > ffi::OMX_MAX_NDIM {
            return Err(Exception::custom("bad shape/strides"));
        }
        let mut desc = ffi::omx_array { data, dtype: dtype.code(), ndim: shape.len() as i32,
                                        shape: [0; ffi::OMX_MAX_NDIM], strides: [0; ffi::OMX_MAX_NDIM] };
        desc.shape[..shape.len()].copy_from_slice(shape);
        desc.strides[..strides.len()].copy_from_slice(strides);
        Ok(Self { desc, _storage: Arc::new(Storage::External) })
    }

    /// Fresh contiguous device array (uninitialised).
    pub fn empty(shape: &[i64], dtype: Dtype) -> Result<Self> {
        let n: i64 = shape.iter().product();
        let mut ptr = std::ptr::null_mut();
        let rc = unsafe { cudaMalloc(&mut ptr, (n.max(1) as usize) * dtype.size()) };
        if rc != 0 {
            return Err(Exception::custom(format!("cudaMalloc failed with {rc}")));
        }
        let mut strides = vec![1i64; shape.len()];
        for i in (0..shape.len().saturating_sub(1)).rev() {
            strides[i] = strides[i + 1] * shape[i + 1];
        }
        let mut a = unsafe { Self::from_raw(ptr, dtype, shape, &strides)? };
        a._storage = Arc::new(Storage::Owned { ptr, free: cuda_free });
        Ok(a)
    }

    pub(crate) fn from_desc(desc: ffi::omx_array, keep: Arc<dyn std::any::Any + Send + Sync>) -> Self {
        Self { desc, _storage: Arc::new(Storage::Borrowed(keep)) }
    }

    pub fn shape(&self) -> &[i64] {
        &self.desc.shape[..self.desc.ndim as usize]
    }
    pub fn strides(&self) -> &[i64] {
        &self.desc.strides[..self.desc.ndim as usize]
    }
    pub fn dtype(&self) -> Dtype {
        Dtype::from_code(self.desc.dtype)
    }
    pub fn ndim(&self) -> usize {
        self.desc.ndim as usize
    }
    pub fn as_ptr(&self) -> *const ffi::omx_array {
        &self.desc
    }

    /// `transpose_axes(&[0, 2, 1, 3])`-style view: permutes shape and strides, no copy.
    pub fn transpose_axes(&self, axes: &[usize]) -> Result<Self> {
        if axes.len() != self.ndim() {
            return Err(Exception::custom("transpose_axes: wrong number of axes"));
        }
        let mut out = self.clone();
        for (i, &a) in axes.iter().enumerate() {
            out.desc.shape[i] = self.desc.shape[a];
            out.desc.strides[i] = self.desc.strides[a];
        }
        Ok(out)
    }
}

impl AsRef<Array> for Array {
    fn as_ref(&self) -> &Array {
        self
    }
}

pub(crate) fn check(status: i32) -> Result<()> {
    guard(status)
}
