//! rsrl_b200_shim — rsrl's trait surface over the C ABI of `include/rsrl_b200.h`.
//!
//! NOT COMPILED in this repository's build image (no cargo/rustc); it documents, as code, the
//! reference-side binding for the hot path:
//!   `Domain::transition`          rsrl_domains/src/lib.rs:436-446      -> `rsrl_domain_step` / fused in `rsrl_engine_step`
//!   `Function<(S,)>::evaluate`    rsrl/src/fa/linear.rs:303-311        -> `rsrl_engine_evaluate`
//!   `Handler<&Transition>::handle` rsrl/src/control/td/q_learning.rs:51-71 (sarsa.rs:53-75, expected_sarsa.rs:45-66)
//!                                                                      -> `rsrl_engine_handle`
//!   `Policy::sample` / `mode`     rsrl/src/policies/mod.rs:65-78       -> `rsrl_engine_sample` / `rsrl_engine_mode`
//!   `Parameterised::weights`      rsrl/src/params/mod.rs:116-134       -> `rsrl_engine_get_weights`
//! and, for N >> 1, the batched loop of examples/q_learning.rs:34-55    -> `rsrl_engine_step(k)`.
#![allow(non_camel_case_types)]
use ndarray::Array2;
use rsrl::domains::{Observation, Transition};
use rsrl::{params::Parameterised, policies::Policy, Handler};
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rsrl_config_t {
    pub struct_size: u32,
    pub domain: i32,
    pub basis: i32,
    pub basis_order: i32,
    pub n_tilings: i32,
    pub tiles_per_dim: i32,
    pub memory_size: i32,
    pub algo: i32,
    pub policy: i32,
    pub trace_rule: i32,
    pub weight_mode: i32,
    pub update_scale: i32,
    pub dtype: i32,
    pub init_mode: i32,
    pub device: i32,
    pub record_td_error: i32,
    pub n_envs: i64,
    pub env_offset: i64,
    pub n_envs_global: i64,
    pub max_episode_steps: i64,
    pub seed: u64,
    pub lr: f64,
    pub alpha: f64,
    pub gamma: f64,
    pub lambda: f64,
    pub epsilon: f64,
    pub init_lo: [f64; 4],
    pub init_hi: [f64; 4],
}

#[repr(C)]
pub struct rsrl_engine_t {
    _private: [u8; 0],
}

extern "C" {
    pub fn rsrl_last_error() -> *const c_char;
    pub fn rsrl_config_default(cfg: *mut rsrl_config_t) -> c_int;
    pub fn rsrl_engine_create(cfg: *const rsrl_config_t, out: *mut *mut rsrl_engine_t) -> c_int;
    pub fn rsrl_engine_destroy(e: *mut rsrl_engine_t) -> c_int;
    pub fn rsrl_engine_reset(e: *mut rsrl_engine_t, init_states: *const f64) -> c_int;
    pub fn rsrl_engine_step(e: *mut rsrl_engine_t, k_steps: i64) -> c_int;
    pub fn rsrl_engine_sync(e: *mut rsrl_engine_t) -> c_int;
    pub fn rsrl_engine_get_states(e: *mut rsrl_engine_t, out: *mut f64) -> c_int;
    pub fn rsrl_engine_get_weights(e: *mut rsrl_engine_t, out: *mut f64) -> c_int;
    pub fn rsrl_engine_set_weights(e: *mut rsrl_engine_t, w: *const f64) -> c_int;
    pub fn rsrl_engine_evaluate(e: *mut rsrl_engine_t, n: i64, states: *const f64, q_out: *mut f64) -> c_int;
    pub fn rsrl_engine_sample(e: *mut rsrl_engine_t, n: i64, states: *const f64, draw: u64, actions_out: *mut i32) -> c_int;
    pub fn rsrl_engine_mode(e: *mut rsrl_engine_t, n: i64, states: *const f64, actions_out: *mut i32) -> c_int;
    pub fn rsrl_engine_handle(
        e: *mut rsrl_engine_t, n: i64, from_states: *const f64, actions: *const i32, rewards: *const f64,
        to_states: *const f64, terminal: *const u8, draw: u64, td_out: *mut f64,
    ) -> c_int;
    pub fn rsrl_domain_step(
        domain: i32, n: i64, states_inout: *mut f64, actions: *const i32, rewards_out: *mut f64, terminal_out: *mut u8,
    ) -> c_int;
}

#[derive(Debug)]
pub struct Error(pub i32, pub String);

fn check(code: c_int) -> Result<(), Error> {
    if code == 0 {
        Ok(())
    } else {
        let msg = unsafe { std::ffi::CStr::from_ptr(rsrl_last_error()) }.to_string_lossy().into_owned();
        Err(Error(code, msg))
    }
}

/// Owns one `rsrl_engine_t`.  `!Send + !Sync` like the reference's `Shared<T> = Rc<RefCell<T>>` (core.rs:13-15).
pub struct GpuAgent {
    raw: *mut rsrl_engine_t,
    cfg: rsrl_config_t,
    n_features: usize,
    n_actions: usize,
    draws: std::cell::Cell<u64>,
    _not_send: std::marker::PhantomData<std::rc::Rc<c_void>>,
}

impl GpuAgent {
    /// `examples/q_learning.rs:24-32`: Fourier(5)+bias LFA, SGD(0.001), QLearning{gamma: 0.9}, Greedy.
    pub fn q_learning_example() -> Result<Self, Error> {
        let mut cfg: rsrl_config_t = unsafe { std::mem::zeroed() };
        check(unsafe { rsrl_config_default(&mut cfg) })?;
        Self::new(cfg, 36, 3)
    }

    pub fn new(cfg: rsrl_config_t, n_features: usize, n_actions: usize) -> Result<Self, Error> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { rsrl_engine_create(&cfg, &mut raw) })?;
        Ok(GpuAgent { raw, cfg, n_features, n_actions, draws: std::cell::Cell::new(0), _not_send: Default::default() })
    }

    /// The batched fused loop: k iterations of examples/q_learning.rs:40-52 for all `n_envs` envs.
    pub fn step_many(&mut self, k: i64) -> Result<(), Error> {
        check(unsafe { rsrl_engine_step(self.raw, k) })?;
        check(unsafe { rsrl_engine_sync(self.raw) })
    }
}

impl Drop for GpuAgent {
    fn drop(&mut self) {
        unsafe { rsrl_engine_destroy(self.raw) };
    }
}

/// `impl Handler<&Transition<S, usize>> for QLearning<Q>` (control/td/q_learning.rs:42-71), N = 1 view.
impl<'m> Handler<&'m Transition<Vec<f64>, usize>> for GpuAgent {
    type Response = f64; // the TD error (Response{error})
    type Error = Error;

    fn handle(&mut self, t: &'m Transition<Vec<f64>, usize>) -> Result<f64, Error> {
        let (a, r) = (t.action as i32, t.reward);
        let term: u8 = match t.to { Observation::Terminal(_) => 1, _ => 0 };
        let mut td = 0.0f64;
        let draw = self.draws.get();
        self.draws.set(draw + 1);
        check(unsafe {
            rsrl_engine_handle(self.raw, 1, t.from.state().as_ptr(), &a, &r, t.to.state().as_ptr(), &term, draw, &mut td)
        })?;
        Ok(td)
    }
}

/// `impl Policy<&Vec<f64>> for Greedy<Q>` / `EpsilonGreedy<Q>` (policies/greedy.rs:74-84, epsilon_greedy.rs:69-83).
/// The caller's `rng` is not consumed: the engine draws from its counter-based Philox stream.
impl<'s> Policy<&'s Vec<f64>> for GpuAgent {
    type Action = usize;

    fn sample<R: rand::Rng + ?Sized>(&self, _rng: &mut R, s: &'s Vec<f64>) -> usize {
        let mut a = 0i32;
        let draw = self.draws.get();
        self.draws.set(draw + 1);
        check(unsafe { rsrl_engine_sample(self.raw, 1, s.as_ptr(), draw, &mut a) }).expect("rsrl_engine_sample");
        a as usize
    }

    fn mode(&self, s: &'s Vec<f64>) -> usize {
        let mut a = 0i32;
        check(unsafe { rsrl_engine_mode(self.raw, 1, s.as_ptr(), &mut a) }).expect("rsrl_engine_mode");
        a as usize
    }
}

/// `impl Parameterised for VectorLFA` (fa/linear.rs:293-301): weights are F x A, row-major.
impl Parameterised for GpuAgent {
    fn weights(&self) -> Array2<f64> {
        let mut w = Array2::<f64>::zeros((self.n_features, self.n_actions));
        check(unsafe { rsrl_engine_get_weights(self.raw, w.as_mut_ptr()) }).expect("rsrl_engine_get_weights");
        w
    }
    fn weights_view(&self) -> ndarray::ArrayView2<f64> {
        unimplemented!("device-resident weights: use weights() (a copy) — views would need host mirroring")
    }
    fn weights_view_mut(&mut self) -> ndarray::ArrayViewMut2<f64> {
        unimplemented!("use rsrl_engine_set_weights")
    }
    fn weights_dim(&self) -> (usize, usize) { (self.n_features, self.n_actions) }
}
