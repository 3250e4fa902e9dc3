//! `alpha-zero` over the tree-pool and network kernels of libomok_b200: the public items the reference's callers use
//! with their names, argument meaning and error behaviour -- `Agent`, `ActionSamplingMode`, `AgentModel`,
//! `ModelIO`, `EnvTurnMode`, `encode_nn_input`, `encode_nn_targets`, `MCTSExecutor`, `ParallelMCTSExecutor`.
//! The tree of an `Agent` lives in a slot of the device tree pool; `Agent.env` mirrors it after every mutation.  The
//! reference's `Agent.mcts` field (the generic CPU tree) has no counterpart: nothing outside alpha-zero reads it.
//! NOT COMPILED in the build image of this repository (no Rust toolchain): see ../README.md.
use environment::{Environment, GameStatus, Turn};
use omok_b200_sys as sys;
use serde::{Deserialize, Serialize};
use std::path::Path;
use tensorflow::{Scope, Session, Status, Tensor, Variable};
use thiserror::Error;

const CELLS: usize = Environment::BOARD_SIZE * Environment::BOARD_SIZE;

fn call(rc: i32) -> Result<(), Status> {
    Status::check(rc)
}

// ------------------------------------------------------------------------------------------------ encoder.rs
#[derive(Debug, Clone, Copy, PartialEq, Eq, Hash)]
pub enum EnvTurnMode {
    Player,
    Opponent,
}

impl EnvTurnMode {
    fn abi(self) -> i32 {
        match self {
            EnvTurnMode::Player => sys::OMK_TURN_MODE_PLAYER,
            EnvTurnMode::Opponent => sys::OMK_TURN_MODE_OPPONENT,
        }
    }
}

/// reference: alpha-zero/src/encoder.rs:10-46 -- `[n, 9, 9, 3]`, one launch of the encode kernel for all environments
pub fn encode_nn_input<'a>(input_count: usize, env_turn_mode: EnvTurnMode, env_iter: impl Iterator<Item = &'a Environment>) -> Tensor<f32> {
    let side = Environment::BOARD_SIZE as u64;
    let mut input = Tensor::new(&[input_count as u64, side, side, 3]);
    let ids: Vec<i32> = env_iter.take(input_count).map(|env| env.slot()).collect();
    sys::with(|c| call(unsafe { sys::omk_env_encode(c.raw(), ids.as_ptr(), ids.len() as i32, env_turn_mode.abi(), input.as_mut_ptr()) }))
        .expect("omk_env_encode");
    input
}

/// reference: alpha-zero/src/encoder.rs:48-68 (host-side packing, nothing to accelerate)
pub fn encode_nn_targets<'a, const N: usize>(
    input_count: usize,
    pi_iter: impl Iterator<Item = &'a [f32; N]>,
    z_iter: impl Iterator<Item = f32>,
) -> (Tensor<f32>, Tensor<f32>) {
    let side = Environment::BOARD_SIZE as u64;
    let mut policy_target = Tensor::new(&[input_count as u64, side, side]);
    let mut value_target = Tensor::new(&[input_count as u64, 1]);
    for (index, (z, pi)) in z_iter.zip(pi_iter).enumerate() {
        policy_target[index * CELLS..(index + 1) * CELLS].copy_from_slice(&pi[..CELLS]);
        value_target[index] = z;
    }
    (policy_target, value_target)
}

// ------------------------------------------------------------------------------------------------ model_io.rs
#[derive(Error, Debug)]
pub enum ModelIOError {
    #[error("Tensorflow error: {0}")]
    Tensorflow(#[from] Status),
    #[error("IO error: {0}")]
    IO(#[from] std::io::Error),
    #[error("Bincode error: {0}")]
    Bincode(#[from] bincode::Error),
}

/// The checkpoint image (alpha-zero/src/model_io.rs:20-24), bincode 1.x default configuration: the same files as the
/// reference writes (and as omok-ai_b200/model_io.py reads and writes).
#[derive(Serialize, Deserialize)]
pub struct SavedData {
    pub variable_names: Vec<String>,
    pub parameters: Vec<Vec<f32>>,
}

pub struct ModelIO {
    pub variables: Vec<Variable>,
}

impl ModelIO {
    pub fn new(variables: Vec<Variable>, _scope: &mut Scope) -> Result<Self, Status> {
        Ok(ModelIO { variables })
    }

    /// reference: model_io.rs:59-90 -- variables out of the library, one bincode file
    pub fn save(&self, _session: &Session, path: impl AsRef<Path>) -> Result<(), ModelIOError> {
        let mut parameters: Vec<Vec<f32>> = self.variables.iter().map(|v| vec![0f32; v.len()]).collect();
        let pointers: Vec<*mut f32> = parameters.iter_mut().map(|p| p.as_mut_ptr()).collect();
        let lens: Vec<i64> = self.variables.iter().map(|v| v.len() as i64).collect();
        sys::with(|c| call(unsafe { sys::omk_net_get_params(c.raw(), pointers.as_ptr(), lens.as_ptr()) }))?;
        let data = SavedData { variable_names: self.variables.iter().map(|v| v.name().to_string()).collect(), parameters };
        bincode::serialize_into(std::io::BufWriter::new(std::fs::File::create(path)?), &data)?;
        Ok(())
    }

    /// reference: model_io.rs:92-120 -- like the reference, the names in the file are ignored and the parameters are
    /// taken in order; a length mismatch is the library's OMK_ERR_INVALID
    pub fn load(&self, _session: &Session, path: impl AsRef<Path>) -> Result<(), ModelIOError> {
        let data: SavedData = bincode::deserialize_from(std::io::BufReader::new(std::fs::File::open(path)?))?;
        if data.parameters.len() < self.variables.len() {
            return Err(ModelIOError::Tensorflow(Status::from_message("checkpoint holds fewer tensors than the network has variables")));
        }
        let pointers: Vec<*const f32> = data.parameters.iter().take(self.variables.len()).map(|p| p.as_ptr()).collect();
        let lens: Vec<i64> = data.parameters.iter().take(self.variables.len()).map(|p| p.len() as i64).collect();
        sys::with(|c| call(unsafe { sys::omk_net_load_params(c.raw(), pointers.as_ptr(), lens.as_ptr()) }))?;
        Ok(())
    }
}

// ------------------------------------------------------------------------------------------------ agent_model.rs
/// The 31 variables in the reference's graph / checkpoint order with the op names its builders produce
/// (network-utils/src/lib.rs:138,147,203,231,250,305,314; network.rs:66,97,140,153,189,228).
fn network_variables() -> Vec<Variable> {
    let mut v = vec![Variable::new("conv_w", 3 * 128), Variable::new("conv_b", 128)];
    for i in 0..3 {
        v.push(Variable::new(format!("residual_{i}_conv0_w"), 128 * 32));
        v.push(Variable::new(format!("residual_{i}_conv0_b"), 32));
        v.push(Variable::new(format!("residual_{i}_conv1_w"), 3 * 3 * 32));
        v.push(Variable::new(format!("residual_{i}_conv1_w_1"), 32 * 32));
        v.push(Variable::new(format!("residual_{i}_conv1_b"), 32));
        v.push(Variable::new(format!("residual_{i}_conv2_w"), 32 * 128));
        v.push(Variable::new(format!("residual_{i}_conv2_b"), 128));
    }
    v.extend([
        Variable::new("fc0_w", 10368 * 512),
        Variable::new("fc0_b", 512),
        Variable::new("fc1_w", 512 * 512),
        Variable::new("fc1_b", 512),
        Variable::new("v_fc0_w", 512),
        Variable::new("v_fc0_b", 1),
        Variable::new("p_fc0_w", 512 * 81),
        Variable::new("p_fc0_b", 81),
    ]);
    v
}

pub struct AgentModel {
    pub variables: Vec<Variable>,
    pub io: ModelIO,
}

impl AgentModel {
    pub const LEARNING_RATE: f32 = 0.01;

    /// reference: agent_model.rs:26-103.  The graph is fixed inside the library; the weights exist once the callers
    /// have run the initialisers (tensorflow::Session::run) or `io.load`.
    pub fn new(scope: &mut Scope) -> Result<Self, Status> {
        let variables = network_variables();
        let io = ModelIO::new(variables.clone(), scope)?;
        Ok(AgentModel { variables, io })
    }

    fn evaluate(&self, input: &Tensor<f32>, want_v: bool) -> Result<(Tensor<f32>, Tensor<f32>), Status> {
        let n = input.len() / (CELLS * 3);
        let side = Environment::BOARD_SIZE as u64;
        let mut p = Tensor::new(&[n as u64, side, side]);
        let mut v = Tensor::new(&[n as u64, 1]);
        let v_ptr = if want_v { v.as_mut_ptr() } else { std::ptr::null_mut() };
        sys::with(|c| call(unsafe { sys::omk_net_eval_images(c.raw(), input.as_ptr(), n as i32, p.as_mut_ptr(), v_ptr) }))?;
        Ok((p, v))
    }

    /// reference: agent_model.rs:105-114
    pub fn evaluate_p(&self, _session: &Session, input: Tensor<f32>) -> Result<Tensor<f32>, Status> {
        Ok(self.evaluate(&input, false)?.0)
    }

    /// reference: agent_model.rs:116-134
    pub fn evaluate_pv(&self, _session: &Session, input: Tensor<f32>) -> Result<(Tensor<f32>, Tensor<f32>), Status> {
        self.evaluate(&input, true)
    }

    /// reference: agent_model.rs:136-168 -- one Adadelta step on the minibatch, then the second forward that reports
    /// `(p_loss, v_loss, loss)`.  `input` is encode_nn_input's `[n,9,9,3]` tensor, `policy_target` `[n,9,9]`,
    /// `value_target` `[n,1]` (encoder.rs:48-68).  Forwards to omk_train_step; the optimizer slots live in the context.
    pub fn train(&self, _session: &Session, input: Tensor<f32>, policy_target: Tensor<f32>, value_target: Tensor<f32>) -> Result<(f32, f32, f32), Status> {
        let n = input.len() / (CELLS * 3);
        debug_assert_eq!(policy_target.len(), n * CELLS);
        debug_assert_eq!(value_target.len(), n);
        let mut losses = [0f32; 3];
        sys::with(|c| {
            call(unsafe {
                sys::omk_train_step(c.raw(), input.as_ptr(), policy_target.as_ptr(), value_target.as_ptr(), n as i32, losses.as_mut_ptr())
            })
        })?;
        Ok((losses[0], losses[1], losses[2]))
    }
}

// ------------------------------------------------------------------------------------------------ agent.rs
/// A method to sample actions from the policy (agent.rs:236-241).
pub enum ActionSamplingMode {
    /// Selects the action with the highest probability.
    Best,
    /// Selects the action using Boltzmann distribution with the given temperature.
    Boltzmann(f32),
}

pub struct Agent {
    pub env: Environment,
    tree: i32, // slot of the device tree pool
}

impl Agent {
    /// reference: agent.rs:16-35 -- root policy = raw evaluate_p of the empty board
    pub fn new(_agent_model: &AgentModel, _session: &Session) -> Result<Self, Status> {
        let tree = sys::with(|c| c.take_tree());
        sys::with(|c| call(unsafe { sys::omk_pool_new_games(c.raw(), &tree, 1, std::ptr::null(), sys::OMK_EVAL_NET) }))?;
        Ok(Agent { env: Environment::new(), tree })
    }

    /// The tree slot, for the batched executor calls.
    pub fn tree(&self) -> i32 {
        self.tree
    }

    /// Debug check: `self.env` (kept in step by `play_action`) equals the environment stored with the tree's root.
    fn env_matches_tree(&self) -> bool {
        let (mut cells, mut turn, mut legal) = ([0u8; CELLS], 0u8, 0u16);
        let ok = sys::with(|c| call(unsafe { sys::omk_pool_get_env(c.raw(), self.tree, cells.as_mut_ptr(), &mut turn, &mut legal) })).is_ok();
        ok && self.env.legal_move_count == legal
            && (self.env.turn == Turn::Black) == (turn == 0)
            && self.env.board.iter().zip(cells.iter()).all(|(stone, &cell)| *stone as u8 == cell)
    }

    /// reference: agent.rs:43-81 -- visit-count policy of the root, `None` when the root has no visits
    pub fn compute_policy(&self) -> Option<[f32; CELLS]> {
        let (mut policy, mut valid) = ([0f32; CELLS], 0u8);
        sys::with(|c| call(unsafe { sys::omk_pool_policy(c.raw(), &self.tree, 1, policy.as_mut_ptr(), &mut valid) })).ok()?;
        (valid != 0).then_some(policy)
    }

    /// reference: agent.rs:83-137 -- `(action, un-heated policy)`
    pub fn sample_action(&self, mode: ActionSamplingMode) -> Option<(usize, [f32; CELLS])> {
        let (abi_mode, temperature) = match mode {
            ActionSamplingMode::Best => (sys::OMK_SAMPLE_BEST, 1.0f32),
            ActionSamplingMode::Boltzmann(t) => (sys::OMK_SAMPLE_BOLTZMANN, t),
        };
        let (mut action, mut policy) = (-1i32, [0f32; CELLS]);
        sys::with(|c| call(unsafe { sys::omk_pool_sample(c.raw(), &self.tree, 1, &abi_mode, &temperature, &mut action, policy.as_mut_ptr()) })).ok()?;
        (action >= 0).then_some((action as usize, policy))
    }

    /// reference: agent.rs:144-197 -- one evaluate_p with the Opponent encoding, then the missing child is created
    pub fn ensure_action_exists(&mut self, action: usize, _agent_model: &AgentModel, _session: &Session) -> Result<(), Status> {
        let action = action as i32;
        sys::with(|c| call(unsafe { sys::omk_pool_ensure_action(c.raw(), &self.tree, &action, 1, sys::OMK_EVAL_NET) }))
    }

    /// reference: agent.rs:206-232 -- `None` when the action has no child, the root is terminal or the cell is occupied
    pub fn play_action(&mut self, action: usize) -> Option<GameStatus> {
        let (action, mut status) = (action as i32, sys::OMK_NONE);
        sys::with(|c| call(unsafe { sys::omk_pool_play(c.raw(), &self.tree, &action, 1, &mut status) })).ok()?;
        let status = GameStatus::from_abi(status)?;
        self.env.place_stone(action as usize); // the public mirror of the tree's environment (agent.rs:214)
        debug_assert!(self.env_matches_tree());
        Some(status)
    }
}

impl Drop for Agent {
    fn drop(&mut self) {
        sys::with(|c| c.give_tree(self.tree));
    }
}

// ------------------------------------------------------------------------------------------------ executors
fn search(ids: &[i32], count: usize, batch_size: usize, epsilon: f32, alpha: f32) -> Result<(), Status> {
    sys::with(|c| {
        call(unsafe {
            sys::omk_pool_search(c.raw(), ids.as_ptr(), ids.len() as i32, count as i32, batch_size as i32, epsilon, alpha, sys::OMK_EVAL_NET)
        })
    })
}

/// reference: parallel_mcts_executor.rs:13-270.  The rayon pool is gone: every agent's tree is searched by its own warp,
/// all trees share each round's network batch.
pub struct ParallelMCTSExecutor;

impl ParallelMCTSExecutor {
    pub fn new() -> Self {
        ParallelMCTSExecutor
    }

    /// `ceil(count / batch_size)` rounds of {batch_size selections + expansions per agent, one network call over all
    /// requests, apply + backup}; per-agent results do not depend on the other agents.
    #[allow(clippy::too_many_arguments)]
    pub fn execute(&self, count: usize, batch_size: usize, epsilon: f32, alpha: f32, _agent_model: &AgentModel, _session: &Session, agents: &[Agent]) -> Result<(), Status> {
        let ids: Vec<i32> = agents.iter().map(|a| a.tree()).collect();
        search(&ids, count, batch_size, epsilon, alpha)
    }
}

impl Default for ParallelMCTSExecutor {
    fn default() -> Self {
        Self::new()
    }
}

/// reference: mcts_executor.rs:16-255.  The reference runs its rounds concurrently on one tree (a data race on the
/// statistics); here the same rounds run in order, which is `execute` with one agent.
pub struct MCTSExecutor;

impl MCTSExecutor {
    pub fn new() -> Self {
        MCTSExecutor
    }

    #[allow(clippy::too_many_arguments)]
    pub fn run(&self, count: usize, batch_size: usize, epsilon: f32, alpha: f32, _agent_model: &AgentModel, _session: &Session, agent: &Agent) -> Result<(), Status> {
        search(&[agent.tree()], count, batch_size, epsilon, alpha)
    }
}

impl Default for MCTSExecutor {
    fn default() -> Self {
        Self::new()
    }
}
