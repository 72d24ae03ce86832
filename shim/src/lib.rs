//! rsrl_b200_shim — rsrl's trait surface over the C ABI of `include/rsrl_b200.h`.
//!
//! NOT COMPILED in this repository's build image (no cargo/rustc); it documents, as code, the
//! reference-side binding for the hot path (INTEGRATION.md lists the `cargo check` a maintainer runs):
//!   `Domain`                      rsrl_domains/src/lib.rs:417-480       -> `GpuDomain` over `rsrl_domain_info/step/is_terminal`
//!   `Function<(S,)>::evaluate`    rsrl/src/fa/linear.rs:303-311        -> `rsrl_engine_evaluate`
//!   `Enumerable<(S,)>`            rsrl/src/core.rs:70-117              -> default bodies over `evaluate` (like `Shared<LFA>`)
//!   `Handler<&Transition>::handle` rsrl/src/control/td/q_learning.rs:51-71 (sarsa.rs:53-75, expected_sarsa.rs:45-66, ...)
//!                                                                      -> `rsrl_engine_handle`
//!   `Policy::sample` / `mode`     rsrl/src/policies/mod.rs:65-78       -> `rsrl_engine_sample` / `rsrl_engine_mode`
//!   `Function<(S, A)>` / `Function<(S, &A)>` (the supertraits `Policy` demands, policies/mod.rs:65-68;
//!       greedy.rs:46-58, epsilon_greedy.rs:47-59)                      -> `rsrl_engine_evaluate` + `rsrl_policy_probs`
//!   `Parameterised`               rsrl/src/params/mod.rs:116-134       -> host mirror of `rsrl_engine_get/set_weights`
//! and, for N >> 1, the batched loop of examples/q_learning.rs:34-55    -> `rsrl_engine_step(k)`.
#![allow(non_camel_case_types)]
use ndarray::{Array2, ArrayView2, ArrayViewMut2};
use rsrl::domains::{Domain, Observation, Transition};
use rsrl::spaces::{discrete::Ordinal, real::Interval, ProductSpace};
use rsrl::{params::Parameterised, policies::Policy, Enumerable, Function, Handler};
use std::cell::{Cell, UnsafeCell};
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
    pub fn rsrl_config_dims(cfg: *const rsrl_config_t, dim: *mut i32, n_actions: *mut i32, n_features: *mut i64) -> c_int;
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
    pub fn rsrl_engine_rollout(
        e: *mut rsrl_engine_t, n: i64, init_states: *const f64, step_limit: i64, greedy: i32, draw: u64, start_out: *mut f64,
        next_out: *mut f64, actions_out: *mut i32, rewards_out: *mut f64, terminal_out: *mut u8, len_out: *mut i32,
    ) -> c_int;
    pub fn rsrl_policy_probs(policy: i32, epsilon: f64, n: i64, n_actions: i32, q: *const f64, probs_out: *mut f64) -> c_int;
    pub fn rsrl_domain_info(domain: i32, dim: *mut i32, n_actions: *mut i32, lo: *mut f64, hi: *mut f64, start: *mut f64) -> c_int;
    pub fn rsrl_domain_step(
        domain: i32, n: i64, states_inout: *mut f64, actions: *const i32, rewards_out: *mut f64, terminal_out: *mut u8,
    ) -> c_int;
    pub fn rsrl_domain_is_terminal(domain: i32, n: i64, states: *const f64, terminal_out: *mut u8) -> c_int;
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

// ---------------------------------------------------------------------------------------------------------------------
// Domain (rsrl_domains/src/lib.rs:417-480): one env whose `step` runs on the GPU (`rsrl_domain_step`, n = 1).  `transition`
// and `rollout` are the trait's provided methods and work unchanged on top of `emit` / `step`.
// ---------------------------------------------------------------------------------------------------------------------
pub struct GpuDomain {
    domain: i32, // rsrl_domain_t: 0 MountainCar, 1 CartPole, 2 Acrobot
    state: Vec<f64>,
    lo: Vec<f64>,
    hi: Vec<f64>,
    n_actions: usize,
}

impl GpuDomain {
    /// `MountainCar::default()` etc.: the domain's default start state (discrete.rs:68-70, cart_pole.rs:76, acrobot.rs:112).
    pub fn new(domain: i32) -> Result<Self, Error> {
        let (mut dim, mut na) = (0i32, 0i32);
        let (mut lo, mut hi, mut start) = ([0f64; 4], [0f64; 4], [0f64; 4]);
        check(unsafe { rsrl_domain_info(domain, &mut dim, &mut na, lo.as_mut_ptr(), hi.as_mut_ptr(), start.as_mut_ptr()) })?;
        let d = dim as usize;
        Ok(GpuDomain { domain, state: start[..d].to_vec(), lo: lo[..d].to_vec(), hi: hi[..d].to_vec(), n_actions: na as usize })
    }

    pub fn mountain_car() -> Result<Self, Error> { Self::new(0) }

    fn terminal(&self) -> bool {
        let mut t = 0u8;
        check(unsafe { rsrl_domain_is_terminal(self.domain, 1, self.state.as_ptr(), &mut t) }).expect("rsrl_domain_is_terminal");
        t != 0
    }
}

impl Domain for GpuDomain {
    type StateSpace = ProductSpace<Interval>;
    type ActionSpace = Ordinal;

    fn state_space(&self) -> Self::StateSpace {
        self.lo.iter().zip(self.hi.iter()).fold(ProductSpace::empty(), |s, (&l, &h)| s + Interval::bounded(l, h))
    }

    fn action_space(&self) -> Ordinal { Ordinal::new(self.n_actions) }

    fn emit(&self) -> Observation<Vec<f64>> {
        if self.terminal() { Observation::Terminal(self.state.clone()) } else { Observation::Full(self.state.clone()) }
    }

    fn step(&mut self, action: &usize) -> (Observation<Vec<f64>>, f64) {
        let (a, mut r, mut t) = (*action as i32, 0f64, 0u8);
        check(unsafe { rsrl_domain_step(self.domain, 1, self.state.as_mut_ptr(), &a, &mut r, &mut t) }).expect("rsrl_domain_step");
        let obs = if t != 0 { Observation::Terminal(self.state.clone()) } else { Observation::Full(self.state.clone()) };
        (obs, r)
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The agent + its policy + its function approximator: one `rsrl_engine_t`.
// `!Send + !Sync` like the reference's `Shared<T> = Rc<RefCell<T>>` (core.rs:13-15).
// ---------------------------------------------------------------------------------------------------------------------
pub struct GpuAgent {
    raw: *mut rsrl_engine_t,
    cfg: rsrl_config_t,
    n_features: usize,
    n_actions: usize,
    draws: Cell<u64>,
    /// host mirror of the device weights for `Parameterised::weights_view{,_mut}`: refreshed before a view is handed out,
    /// written back (if `weights_view_mut` was taken) before the next device call
    mirror: UnsafeCell<Array2<f64>>,
    mirror_dirty: Cell<bool>,
    _not_send: std::marker::PhantomData<std::rc::Rc<c_void>>,
}

impl GpuAgent {
    /// `examples/q_learning.rs:24-32`: Fourier(5)+bias LFA, SGD(0.001), QLearning{gamma: 0.9}, Greedy.
    pub fn q_learning_example() -> Result<Self, Error> {
        let mut cfg: rsrl_config_t = unsafe { std::mem::zeroed() };
        check(unsafe { rsrl_config_default(&mut cfg) })?;
        Self::new(cfg)
    }

    pub fn new(cfg: rsrl_config_t) -> Result<Self, Error> {
        let (mut d, mut a, mut f) = (0i32, 0i32, 0i64);
        check(unsafe { rsrl_config_dims(&cfg, &mut d, &mut a, &mut f) })?;
        let mut raw = std::ptr::null_mut();
        check(unsafe { rsrl_engine_create(&cfg, &mut raw) })?;
        Ok(GpuAgent {
            raw, cfg, n_features: f as usize, n_actions: a as usize, draws: Cell::new(0),
            mirror: UnsafeCell::new(Array2::zeros((f as usize, a as usize))), mirror_dirty: Cell::new(false),
            _not_send: Default::default(),
        })
    }

    /// weights edited through `weights_view_mut` reach the device before anything reads them there
    fn flush(&self) {
        if self.mirror_dirty.replace(false) {
            let m = unsafe { &*self.mirror.get() };
            check(unsafe { rsrl_engine_set_weights(self.raw, m.as_ptr()) }).expect("rsrl_engine_set_weights");
        }
    }

    fn refresh(&self) -> &mut Array2<f64> {
        self.flush();
        let m = unsafe { &mut *self.mirror.get() };
        check(unsafe { rsrl_engine_get_weights(self.raw, m.as_mut_ptr()) }).expect("rsrl_engine_get_weights");
        m
    }

    /// The batched fused loop: k iterations of examples/q_learning.rs:40-52 for all `n_envs` envs.
    pub fn step_many(&mut self, k: i64) -> Result<(), Error> {
        self.flush();
        check(unsafe { rsrl_engine_step(self.raw, k) })?;
        check(unsafe { rsrl_engine_sync(self.raw) })
    }

    fn next_draw(&self) -> u64 {
        let d = self.draws.get();
        self.draws.set(d + 1);
        d
    }
}

impl Drop for GpuAgent {
    fn drop(&mut self) {
        unsafe { rsrl_engine_destroy(self.raw) };
    }
}

/// `impl Function<(&S,)> for VectorLFA` (fa/linear.rs:303-311): the Q vector.
impl<'s> Function<(&'s Vec<f64>,)> for GpuAgent {
    type Output = Vec<f64>;

    fn evaluate(&self, (s,): (&'s Vec<f64>,)) -> Vec<f64> {
        self.flush();
        let mut q = vec![0f64; self.n_actions];
        check(unsafe { rsrl_engine_evaluate(self.raw, 1, s.as_ptr(), q.as_mut_ptr()) }).expect("rsrl_engine_evaluate");
        q
    }
}

/// `Enumerable` through the default bodies (`evaluate(args)[i]`, `find_max`: core.rs:79-105) — what `Shared<LFA>` gets too.
impl<'s> Enumerable<(&'s Vec<f64>,)> for GpuAgent {}

/// `impl Handler<&Transition<S, usize>> for QLearning<Q>` (control/td/q_learning.rs:42-71) and the other TD agents, N = 1 view.
impl<'m> Handler<&'m Transition<Vec<f64>, usize>> for GpuAgent {
    type Response = f64; // the TD error (Response{error})
    type Error = Error;

    fn handle(&mut self, t: &'m Transition<Vec<f64>, usize>) -> Result<f64, Error> {
        self.flush();
        let (a, r) = (t.action as i32, t.reward);
        let term: u8 = match t.to { Observation::Terminal(_) => 1, _ => 0 };
        let mut td = 0.0f64;
        let draw = self.next_draw();
        check(unsafe {
            rsrl_engine_handle(self.raw, 1, t.from.state().as_ptr(), &a, &r, t.to.state().as_ptr(), &term, draw, &mut td)
        })?;
        Ok(td)
    }
}

/// The probability of one action: the two `Function` supertraits `Policy<S>` requires (policies/mod.rs:65-68), as
/// `Greedy` / `EpsilonGreedy` / `Softmax` implement them through `self.evaluate((s,))[a]` (greedy.rs:46-58).
fn action_probability(agent: &GpuAgent, s: &Vec<f64>, a: usize) -> f64 {
    let q = Function::<(&Vec<f64>,)>::evaluate(agent, (s,));
    let mut p = vec![0f64; q.len()];
    check(unsafe { rsrl_policy_probs(agent.cfg.policy, agent.cfg.epsilon, 1, q.len() as i32, q.as_ptr(), p.as_mut_ptr()) })
        .expect("rsrl_policy_probs");
    p[a]
}

impl<'s> Function<(&'s Vec<f64>, usize)> for GpuAgent {
    type Output = f64;

    fn evaluate(&self, (s, a): (&'s Vec<f64>, usize)) -> f64 { action_probability(self, s, a) }
}

impl<'s, 'a> Function<(&'s Vec<f64>, &'a usize)> for GpuAgent {
    type Output = f64;

    fn evaluate(&self, (s, a): (&'s Vec<f64>, &'a usize)) -> f64 { action_probability(self, s, *a) }
}

/// `impl Policy<&Vec<f64>> for Greedy<Q>` / `EpsilonGreedy<Q>` / `Softmax<F>` (policies/greedy.rs:74-84, epsilon_greedy.rs:69-83,
/// softmax.rs:131-143).  The caller's `rng` is not consumed: the engine draws from its counter-based Philox stream.
impl<'s> Policy<&'s Vec<f64>> for GpuAgent {
    type Action = usize;

    fn sample<R: rand::Rng + ?Sized>(&self, _rng: &mut R, s: &'s Vec<f64>) -> usize {
        self.flush();
        let mut a = 0i32;
        check(unsafe { rsrl_engine_sample(self.raw, 1, s.as_ptr(), self.next_draw(), &mut a) }).expect("rsrl_engine_sample");
        a as usize
    }

    fn mode(&self, s: &'s Vec<f64>) -> usize {
        self.flush();
        let mut a = 0i32;
        check(unsafe { rsrl_engine_mode(self.raw, 1, s.as_ptr(), &mut a) }).expect("rsrl_engine_mode");
        a as usize
    }
}

/// `impl Parameterised for VectorLFA` (fa/linear.rs:293-301): weights are F x A, row-major — the layout
/// `rsrl_engine_get_weights` returns.  Views go through the host mirror (same lifetime trick as
/// `impl Parameterised for Shared<F>`, params/mod.rs:136-148, which also hands out a view past a `RefCell`).
impl Parameterised for GpuAgent {
    fn weights(&self) -> Array2<f64> { self.refresh().clone() }

    fn weights_view(&self) -> ArrayView2<f64> { self.refresh().view() }

    fn weights_view_mut(&mut self) -> ArrayViewMut2<f64> {
        let m = self.refresh();
        self.mirror_dirty.set(true); // written back by the next call that touches the device
        m.view_mut()
    }

    fn weights_dim(&self) -> (usize, usize) { (self.n_features, self.n_actions) }
}
