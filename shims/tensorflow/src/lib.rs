//! Facade for the `tensorflow` 0.21 items that appear in the public signatures of `alpha-zero` and in its callers
//! (src/trainer.rs:38-48, benchmark/src/agent.rs:18-20, gui/src/agent.rs:16-18): thin handles onto the libomok_b200
//! context, so that those callers compile unchanged.  Wire with `[patch.crates-io] tensorflow = { path = "shims/tensorflow" }`.
//! NOT COMPILED in the build image of this repository (no Rust toolchain): see ../README.md.
use std::fmt;
use std::ops::{Deref, DerefMut, Index, IndexMut};

/// Error of a failed library call (`tensorflow::Status` in the reference's `Result<_, Status>` signatures).
#[derive(Debug, Clone)]
pub struct Status {
    message: String,
}

impl Status {
    pub fn new() -> Self {
        Status { message: String::new() }
    }
    pub fn from_message(message: impl Into<String>) -> Self {
        Status { message: message.into() }
    }
    /// `Err(Status)` carrying omk_last_error() when `rc` is not OMK_OK.
    pub fn check(rc: i32) -> Result<(), Status> {
        omok_b200_sys::check(rc).map_err(Status::from_message)
    }
}

impl Default for Status {
    fn default() -> Self {
        Self::new()
    }
}

impl fmt::Display for Status {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        f.write_str(&self.message)
    }
}

impl std::error::Error for Status {}

/// Dense row-major tensor: the only element type on this path is `f32`.
#[derive(Debug, Clone, PartialEq)]
pub struct Tensor<T> {
    dims: Vec<u64>,
    data: Vec<T>,
}

impl<T: Clone + Default> Tensor<T> {
    pub fn new(dims: &[u64]) -> Self {
        let len = dims.iter().product::<u64>() as usize;
        Tensor { dims: dims.to_vec(), data: vec![T::default(); len] }
    }
    pub fn with_values(mut self, values: &[T]) -> Result<Self, Status> {
        if values.len() != self.data.len() {
            return Err(Status::from_message("Tensor::with_values: length mismatch"));
        }
        self.data.clone_from_slice(values);
        Ok(self)
    }
    pub fn dims(&self) -> &[u64] {
        &self.dims
    }
}

impl<T> Deref for Tensor<T> {
    type Target = [T];
    fn deref(&self) -> &[T] {
        &self.data
    }
}

impl<T> DerefMut for Tensor<T> {
    fn deref_mut(&mut self) -> &mut [T] {
        &mut self.data
    }
}

impl<T, I: std::slice::SliceIndex<[T]>> Index<I> for Tensor<T> {
    type Output = I::Output;
    fn index(&self, index: I) -> &I::Output {
        &self.data[index]
    }
}

impl<T, I: std::slice::SliceIndex<[T]>> IndexMut<I> for Tensor<T> {
    fn index_mut(&mut self, index: I) -> &mut I::Output {
        &mut self.data[index]
    }
}

/// Graph-building scope of the reference (`Scope::new_root_scope()`): the network is fixed inside the library.
#[derive(Debug, Default)]
pub struct Scope;

impl Scope {
    pub fn new_root_scope() -> Self {
        Scope
    }
    pub fn graph(&self) -> Graph {
        Graph
    }
}

#[derive(Debug, Default, Clone, Copy)]
pub struct Graph;

#[derive(Debug, Default)]
pub struct SessionOptions;

impl SessionOptions {
    pub fn new() -> Self {
        SessionOptions
    }
}

/// `Session::new(&SessionOptions::new(), &scope.graph())`: a unit handle; the state lives in the library context.
#[derive(Debug)]
pub struct Session;

impl Session {
    pub fn new(_options: &SessionOptions, _graph: &Graph) -> Result<Self, Status> {
        omok_b200_sys::with(|_| ()); // creates the context (and fails loudly without a B200) at the same point as the reference
        Ok(Session)
    }
}

/// A graph operation the callers only ever pass back to the session: variable initialisers (src/trainer.rs:42-48).
#[derive(Debug, Clone, PartialEq, Eq)]
pub enum Operation {
    InitializeVariable,
    Other,
}

/// One of the 31 network variables (alpha-zero/src/network.rs:78-241), by name and element count; the values live in
/// the library (omk_net_get_params / omk_net_load_params).
#[derive(Debug, Clone)]
pub struct Variable {
    name: String,
    len: usize,
}

impl Variable {
    pub fn new(name: impl Into<String>, len: usize) -> Self {
        Variable { name: name.into(), len }
    }
    pub fn name(&self) -> &str {
        &self.name
    }
    pub fn len(&self) -> usize {
        self.len
    }
    pub fn is_empty(&self) -> bool {
        self.len == 0
    }
    pub fn initializer(&self) -> Operation {
        Operation::InitializeVariable
    }
}

/// Only what src/trainer.rs:42-48 does with it: collect the variable initialisers and run them once.
#[derive(Debug, Default)]
pub struct SessionRunArgs {
    initialise: bool,
}

impl SessionRunArgs {
    pub fn new() -> Self {
        SessionRunArgs { initialise: false }
    }
    pub fn add_target(&mut self, operation: &Operation) {
        self.initialise |= *operation == Operation::InitializeVariable;
    }
}

impl Session {
    /// Running the variable initialisers == the reference's random-init recipe on the device
    /// (network-utils/src/lib.rs:86-92 -> omk_net_init_random; seed from OMOK_B200_SEED, default 0).
    pub fn run(&self, args: &mut SessionRunArgs) -> Result<(), Status> {
        if args.initialise {
            let seed = std::env::var("OMOK_B200_SEED").ok().and_then(|v| v.parse().ok()).unwrap_or(0u64);
            Status::check(omok_b200_sys::with(|c| unsafe { omok_b200_sys::omk_net_init_random(c.raw(), seed) }))?;
        }
        Ok(())
    }
}
