//! `Exception { what }` and the status -> Result guard (mlx-rs/src/error.rs:236-288,
//! mlx-rs/src/utils/guard.rs:24-48).  The library keeps the message in a thread-local slot, so no
//! handler has to be installed; `omx_set_error_handler` is available for hosts that want one.
use std::ffi::CStr;

#[derive(Debug, Clone, thiserror::Error)]
#[error("{what}")]
pub struct Exception {
    pub what: String,
}

impl Exception {
    pub fn custom(what: impl Into<String>) -> Self {
        Self { what: what.into() }
    }
}

pub type Result<T> = std::result::Result<T, Exception>;

#[track_caller]
pub(crate) fn guard(status: i32) -> Result<()> {
    if status == 0 {
        return Ok(());
    }
    let what = unsafe { CStr::from_ptr(crate::ffi::omx_last_error()) }.to_string_lossy().into_owned();
    Err(Exception { what })
}
