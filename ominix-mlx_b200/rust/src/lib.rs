//! omx-attn: drop-in for the attention hot path of mlx-rs / mlx-rs-core on B200.
//!
//! * [`fast::rope`], [`fast::scaled_dot_product_attention`]  <- mlx-rs/src/fast.rs:15-46, 110-151
//! * [`cache::KeyValueCache`], [`cache::KVCache`], [`cache::ConcatKeyValueCache`]
//!                                                          <- mlx-rs-core/src/cache.rs
//! * [`utils`]                                              <- mlx-rs-core/src/utils.rs
//!
//! `Array` here is a strided view of CUDA device memory (the role `mlx_rs::Array` plays in the
//! reference); allocation is delegated to an [`array::DeviceAllocator`] supplied by the host
//! application (cudaMallocAsync by default), because the reference's lazy graph/allocator is out of
//! scope of this path.
pub mod array;
pub mod cache;
pub mod error;
pub mod fast;
pub mod ffi;
pub mod utils;

pub use array::{Array, Dtype, Stream};
pub use error::Exception;
