//! `CudaNBodyPropagator<D>`: the B200 engine behind the reference's own propagator traits.
//!
//! Drop-in for `ephemeris_explorer::dynamics::celestial::NBodyPropagator<D>`
//! (ephemeris_explorer/src/dynamics/celestial.rs:139-140): it satisfies everything
//! `PredictionTarget::Propagator` asks for (ephemeris_explorer/src/prediction.rs:31-37):
//! `Propagator<Solution = Vec<UniformSpline<DVec3>>> + IncrementalPropagator + DirectionalPropagator + Clone + Send + Sync`.
//!
//! SOURCE ONLY -- never compiled in the build image (no Rust toolchain there).
use ephemeris::{
    DirectionalPropagator, IncrementalPropagator, Polynomial, PropagationDirection, Propagator,
    UniformSpline,
};
use ftime::{Duration, Epoch};
use glam::DVec3;
use std::os::raw::{c_char, c_void};

#[repr(C)]
pub struct EeNBody {
    _private: [u8; 0],
}

unsafe extern "C" {
    fn ee_last_error() -> *const c_char;
    fn ee_nbody_create(
        n: i64,
        positions: *const f64,
        velocities: *const f64,
        mus: *const f64,
        t0: f64,
        h_signed: f64,
        method: i32,
        mode: i32,
        device: i32,
        out: *mut *mut EeNBody,
    ) -> i32;
    fn ee_nbody_set_solout(h: *mut EeNBody, delta: f64, periods: *const f64, degrees: *const i32) -> i32;
    fn ee_nbody_step(h: *mut EeNBody, n_steps: i64) -> i32;
    fn ee_nbody_solution_time(h: *mut EeNBody, epoch: *mut f64) -> i32;
    fn ee_nbody_has_reached(h: *mut EeNBody, epoch: f64, reached: *mut i32) -> i32;
    fn ee_nbody_solution_sizes(h: *mut EeNBody, n_poly: *mut i64) -> i32;
    fn ee_nbody_take_solution(
        h: *mut EeNBody,
        start: *mut f64,
        interval: *mut f64,
        coeffs: *mut f64,
        n_coef: *mut i32,
    ) -> i32;
    fn ee_nbody_clone(h: *mut EeNBody, out: *mut *mut EeNBody) -> i32;
    fn ee_nbody_destroy(h: *mut EeNBody);
}

/// Mirrors `ephemeris::NBodyPropagatorError` (nbody.rs:43-47) plus engine failures.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum CudaPropagatorError {
    StepSizeUnderflow,
    MaxIterationsReached,
    BoundReached,
    EvalFailed,
    Solout,
    Engine(i32),
}
impl std::fmt::Display for CudaPropagatorError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "{self:?}")
    }
}
impl std::error::Error for CudaPropagatorError {}

fn status(code: i32) -> Result<(), CudaPropagatorError> {
    match code {
        0 => Ok(()),
        1 => Err(CudaPropagatorError::StepSizeUnderflow),
        2 => Err(CudaPropagatorError::MaxIterationsReached),
        3 => Err(CudaPropagatorError::BoundReached),
        4 => Err(CudaPropagatorError::EvalFailed),
        5 => Err(CudaPropagatorError::Solout),
        c => Err(CudaPropagatorError::Engine(c)),
    }
}

pub struct CudaNBodyPropagator<D> {
    handle: *mut EeNBody,
    n: usize,
    _marker: std::marker::PhantomData<D>,
}

// One stepping thread at a time, exactly how prediction.rs:385-391 uses a propagator.
unsafe impl<D> Send for CudaNBodyPropagator<D> {}
unsafe impl<D> Sync for CudaNBodyPropagator<D> {}

impl<D: PropagationDirection> CudaNBodyPropagator<D> {
    /// Same arguments as `NBodyPropagator::new` (nbody.rs:93-100) with the `SplineInterpolators` solout flattened
    /// into `(delta, sample_period, degree)` per body (dynamics/celestial.rs:156-186).
    pub fn new(
        direction: D,
        initial_time: Epoch,
        positions: Vec<DVec3>,
        velocities: Vec<DVec3>,
        gravitational_parameters: Vec<f64>,
        delta: Duration,
        sample_periods: Vec<Duration>,
        degrees: Vec<i32>,
    ) -> Result<Self, CudaPropagatorError> {
        let n = positions.len();
        let periods: Vec<f64> = sample_periods.iter().map(|d| d.as_seconds()).collect();
        let mut handle = std::ptr::null_mut();
        // DVec3 is repr(C) {x, y, z}: Vec<DVec3> is already the AoS double[3] layout the C ABI takes.
        status(unsafe {
            ee_nbody_create(
                n as i64,
                positions.as_ptr() as *const f64,
                velocities.as_ptr() as *const f64,
                gravitational_parameters.as_ptr(),
                initial_time.as_offset_seconds(),
                direction.signed_delta().as_seconds(),
                12, // QuinlanTremaine12
                0,  // EE_MODE_PARITY: bit-for-bit the CPU path
                0,
                &mut handle,
            )
        })?;
        status(unsafe { ee_nbody_set_solout(handle, delta.as_seconds(), periods.as_ptr(), degrees.as_ptr()) })?;
        Ok(Self { handle, n, _marker: std::marker::PhantomData })
    }
}

impl<D> Drop for CudaNBodyPropagator<D> {
    fn drop(&mut self) {
        unsafe { ee_nbody_destroy(self.handle) }
    }
}

impl<D> Clone for CudaNBodyPropagator<D> {
    fn clone(&self) -> Self {
        let mut out = std::ptr::null_mut();
        let code = unsafe { ee_nbody_clone(self.handle, &mut out) };
        assert_eq!(code, 0, "ee_nbody_clone failed: {:?}", unsafe { std::ffi::CStr::from_ptr(ee_last_error()) });
        Self { handle: out, n: self.n, _marker: std::marker::PhantomData }
    }
}

impl<D> Propagator for CudaNBodyPropagator<D> {
    type Solution = Vec<UniformSpline<DVec3>>;

    fn take_solution(&mut self) -> Self::Solution {
        let mut n_poly = vec![0i64; self.n];
        unsafe { ee_nbody_solution_sizes(self.handle, n_poly.as_mut_ptr()) };
        let total: usize = n_poly.iter().map(|&c| c as usize).sum();
        let (mut start, mut interval) = (vec![0.0; self.n], vec![0.0; self.n]);
        let mut coeffs = vec![0.0f64; total.max(1) * 27];
        let mut n_coef = vec![0i32; total.max(1)];
        unsafe {
            ee_nbody_take_solution(
                self.handle,
                start.as_mut_ptr(),
                interval.as_mut_ptr(),
                coeffs.as_mut_ptr(),
                n_coef.as_mut_ptr(),
            )
        };
        let mut k = 0;
        (0..self.n)
            .map(|b| {
                let mut spline =
                    UniformSpline::new(Epoch::from_offset_seconds(start[b]), Duration::from_seconds(interval[b]));
                for _ in 0..n_poly[b] {
                    let c = &coeffs[k * 27..];
                    let poly = (0..n_coef[k] as usize)
                        .map(|i| DVec3::new(c[3 * i], c[3 * i + 1], c[3 * i + 2]))
                        .collect::<smallvec::SmallVec<[DVec3; 8]>>();
                    spline.push_back(Polynomial::new(poly));
                    k += 1;
                }
                spline
            })
            .collect()
    }
}

impl<D> IncrementalPropagator for CudaNBodyPropagator<D> {
    type Error = CudaPropagatorError;

    fn step(&mut self) -> Result<(), Self::Error> {
        status(unsafe { ee_nbody_step(self.handle, 1) })
    }
}

impl<D: PropagationDirection> DirectionalPropagator for CudaNBodyPropagator<D> {
    fn offset(to: Epoch, duration: Duration) -> Epoch {
        D::offset(to, duration)
    }

    fn distance(from: Epoch, to: Epoch) -> Duration {
        D::distance(from, to)
    }

    fn time(&self) -> Epoch {
        let mut t = 0.0;
        unsafe { ee_nbody_solution_time(self.handle, &mut t) };
        Epoch::from_offset_seconds(t)
    }

    fn has_reached(&self, time: Epoch) -> bool {
        let mut r = 0;
        unsafe { ee_nbody_has_reached(self.handle, time.as_offset_seconds(), &mut r) };
        r != 0
    }
}

#[allow(dead_code)]
fn _assert_bounds<D: PropagationDirection + 'static>() {
    fn needs<T: Propagator + IncrementalPropagator + DirectionalPropagator + Clone + Send + Sync>() {}
    needs::<CudaNBodyPropagator<D>>();
    let _ = std::mem::size_of::<*mut c_void>();
}
