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

// ---------------------------------------------------------------------------------------------------------
// Ships: a batch of `ephemeris::SpacecraftPropagator<[StateVector;1], ReferenceFrame, Bodies, Verner87, CubicHermiteSplineSolout>`
// (ephemeris/src/propagators/spacecraft.rs:415-643) over the device-resident spline ephemeris.
#[repr(C)]
pub struct EeEphem {
    _private: [u8; 0],
}
#[repr(C)]
pub struct EeShips {
    _private: [u8; 0],
}

/// `AdaptiveMethodParams<f64, AbsTol, f64>` (integration/src/lib.rs:174-197) in the C layout of `ee_adaptive_params`.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct EeAdaptiveParams {
    pub h_init: f64,
    pub h_max: f64,
    pub tol_position: f64,
    pub tol_velocity: f64,
    pub fac_min: f64,
    pub fac_max: f64,
    pub fac: f64,
    pub n_max: u32,
    /// 0 = glibc's pow (what `f64::powf` is on Linux, the reference as built), 1 = correctly rounded
    pub pow_mode: u32,
    /// `IntegrationMethod` (flight_plan.rs:175-184): 0 Verner87, 1 CashKarp45, 2 DormandPrince54, 3 DormandPrince87,
    /// 4 Fehlberg45, 5 Tsitouras75, 6 Verner98, 7 Fine45
    pub method: u32,
}

unsafe extern "C" {
    fn ee_nbody_take_solution_ephem(h: *mut EeNBody, out: *mut *mut EeEphem) -> i32;
    fn ee_ephem_destroy(e: *mut EeEphem);
    fn ee_ships_create(
        ephem: *mut EeEphem,
        n_ships: i64,
        t0: *const f64,
        states: *const f64,
        params: *const EeAdaptiveParams,
        burn_offsets: *const i64,
        burn_start: *const f64,
        burn_end: *const f64,
        burn_acc: *const f64,
        burn_ref: *const i32,
        out: *mut *mut EeShips,
    ) -> i32;
    fn ee_ships_step_to(h: *mut EeShips, t_end: f64, max_steps: i64) -> i32;
    fn ee_ships_info(
        h: *mut EeShips,
        status: *mut i32,
        time: *mut f64,
        n_knots: *mut i64,
        n_attempts: *mut u32,
        rhs_evals: *mut u64,
    ) -> i32;
    fn ee_ships_take_knots(h: *mut EeShips, knot_offsets: *const i64, knots7: *mut f64) -> i32;
    fn ee_ships_enable_analytics(h: *mut EeShips, soi_radius: *const f64) -> i32;
    fn ee_ships_analytics_counts(h: *mut EeShips, n_transitions: *mut i32, n_apsides: *mut i32) -> i32;
    fn ee_ships_read_analytics(
        h: *mut EeShips,
        transition_offsets: *const i64,
        transition_time: *mut f64,
        transition_body: *mut i32,
        apsis_offsets: *const i64,
        apsis_time: *mut f64,
        apsis_distance: *mut f64,
        apsis_body: *mut i32,
        apsis_kind: *mut i32,
    ) -> i32;
    fn ee_ships_evaluate_relative(
        h: *mut EeShips,
        ship: i64,
        reference: i32,
        n_times: i64,
        times: *const f64,
        pos: *mut f64,
        vel: *mut f64,
        ok: *mut i32,
    ) -> i32;
    fn ee_ships_destroy(h: *mut EeShips);
}

/// One burn of a `Timeline` (spacecraft.rs:59-70): `reference` = index of the body whose TNB frame the thrust is given
/// in (dynamics/spacecraft.rs:254-293), `None` = inertial.
pub struct ShimBurn {
    pub start: Epoch,
    pub end: Epoch,
    pub acceleration: DVec3,
    pub reference: Option<usize>,
}

/// Device-resident `Vec<UniformSpline<DVec3>>` taken straight from a `CudaNBodyPropagator` (no host round trip).
pub struct CudaEphemeris(*mut EeEphem);
unsafe impl Send for CudaEphemeris {}
unsafe impl Sync for CudaEphemeris {}
impl Drop for CudaEphemeris {
    fn drop(&mut self) {
        unsafe { ee_ephem_destroy(self.0) }
    }
}
impl<D> CudaNBodyPropagator<D> {
    pub fn take_solution_ephemeris(&mut self) -> Result<CudaEphemeris, CudaPropagatorError> {
        let mut e = std::ptr::null_mut();
        status(unsafe { ee_nbody_take_solution_ephem(self.handle, &mut e) })?;
        Ok(CudaEphemeris(e))
    }
}

pub struct CudaSpacecraftBatch<'a> {
    handle: *mut EeShips,
    n: usize,
    _context: &'a CudaEphemeris,
}

impl<'a> CudaSpacecraftBatch<'a> {
    /// n x `SpacecraftPropagator::new(initial_time, initial_state, params, timeline, context, solout)` (spacecraft.rs:453-477).
    pub fn new(
        context: &'a CudaEphemeris,
        initial_times: &[Epoch],
        initial_states: &[(DVec3, DVec3)],
        params: EeAdaptiveParams,
        timelines: &[Vec<ShimBurn>],
    ) -> Result<Self, CudaPropagatorError> {
        let n = initial_states.len();
        let t0: Vec<f64> = initial_times.iter().map(|t| t.as_offset_seconds()).collect();
        let states: Vec<f64> = initial_states
            .iter()
            .flat_map(|(p, v)| [p.x, p.y, p.z, v.x, v.y, v.z])
            .collect();
        let (mut off, mut bs, mut be, mut ba, mut br) = (vec![0i64], vec![], vec![], vec![], vec![]);
        for tl in timelines {
            for b in tl {
                bs.push(b.start.as_offset_seconds());
                be.push(b.end.as_offset_seconds());
                ba.extend_from_slice(&[b.acceleration.x, b.acceleration.y, b.acceleration.z]);
                br.push(b.reference.map_or(-1, |r| r as i32));
            }
            off.push(bs.len() as i64);
        }
        let mut handle = std::ptr::null_mut();
        status(unsafe {
            ee_ships_create(
                context.0,
                n as i64,
                t0.as_ptr(),
                states.as_ptr(),
                &params,
                off.as_ptr(),
                bs.as_ptr(),
                be.as_ptr(),
                ba.as_ptr(),
                br.as_ptr(),
                &mut handle,
            )
        })?;
        Ok(Self { handle, n, _context: context })
    }

    /// Every ship: `IncrementalPropagator::step_to(time)` (ephemeris/src/lib.rs:47-58).
    pub fn step_to(&mut self, time: Epoch, max_steps: i64) -> Result<(), CudaPropagatorError> {
        status(unsafe { ee_ships_step_to(self.handle, time.as_offset_seconds(), max_steps) })
    }

    /// Per ship: `Propagator::take_solution` of `CubicHermiteSplineSolout` as (t, position, velocity) knots, plus the
    /// per-ship status (`StepError` codes; a ship that left the ephemeris reports EvalFailed and keeps its knots).
    pub fn take_solution(&mut self) -> (Vec<Vec<(Epoch, DVec3, DVec3)>>, Vec<i32>) {
        let mut st = vec![0i32; self.n];
        let mut nk = vec![0i64; self.n];
        unsafe {
            ee_ships_info(
                self.handle,
                st.as_mut_ptr(),
                std::ptr::null_mut(),
                nk.as_mut_ptr(),
                std::ptr::null_mut(),
                std::ptr::null_mut(),
            )
        };
        let mut off = vec![0i64; self.n + 1];
        for i in 0..self.n {
            off[i + 1] = off[i] + nk[i];
        }
        let mut flat = vec![0.0f64; off[self.n] as usize * 7];
        unsafe { ee_ships_take_knots(self.handle, off.as_ptr(), flat.as_mut_ptr()) };
        let sol = (0..self.n)
            .map(|i| {
                (off[i] as usize..off[i + 1] as usize)
                    .map(|k| {
                        let c = &flat[k * 7..k * 7 + 7];
                        (
                            Epoch::from_offset_seconds(c[0]),
                            DVec3::new(c[1], c[2], c[3]),
                            DVec3::new(c[4], c[5], c[6]),
                        )
                    })
                    .collect()
            })
            .collect();
        (sol, st)
    }
}

/// `SoiTransitions` / `Apsides` of one ship (dynamics/spacecraft.rs:302-440) as body indices in construction order.
pub struct ShipAnalytics {
    pub transitions: Vec<(Epoch, usize)>,
    /// (time, distance, body, is_apoapsis)
    pub apsides: Vec<(Epoch, f64, usize, bool)>,
}

impl CudaSpacecraftBatch<'_> {
    /// Switch the solution to `SpacecraftSolout`'s (dynamics/spacecraft.rs:448-586); `soi_radius[b]` = `SphereOfInfluence::radius`.
    pub fn enable_analytics(&mut self, soi_radius: &[f64]) -> Result<(), CudaPropagatorError> {
        status(unsafe { ee_ships_enable_analytics(self.handle, soi_radius.as_ptr()) })
    }

    /// Transitions and apsides found since the last `take_solution` (read before taking the solution).
    pub fn analytics(&mut self) -> Result<Vec<ShipAnalytics>, CudaPropagatorError> {
        let (mut ntr, mut nap) = (vec![0i32; self.n], vec![0i32; self.n]);
        status(unsafe { ee_ships_analytics_counts(self.handle, ntr.as_mut_ptr(), nap.as_mut_ptr()) })?;
        let (mut to, mut ao) = (vec![0i64; self.n + 1], vec![0i64; self.n + 1]);
        for i in 0..self.n {
            to[i + 1] = to[i] + ntr[i] as i64;
            ao[i + 1] = ao[i] + nap[i] as i64;
        }
        let (nt, na) = (to[self.n] as usize, ao[self.n] as usize);
        let (mut tt, mut tb) = (vec![0.0f64; nt.max(1)], vec![0i32; nt.max(1)]);
        let (mut at, mut ad, mut ab, mut ak) = (vec![0.0f64; na.max(1)], vec![0.0f64; na.max(1)], vec![0i32; na.max(1)], vec![0i32; na.max(1)]);
        status(unsafe {
            ee_ships_read_analytics(
                self.handle,
                to.as_ptr(),
                tt.as_mut_ptr(),
                tb.as_mut_ptr(),
                ao.as_ptr(),
                at.as_mut_ptr(),
                ad.as_mut_ptr(),
                ab.as_mut_ptr(),
                ak.as_mut_ptr(),
            )
        })?;
        Ok((0..self.n)
            .map(|i| ShipAnalytics {
                transitions: (to[i] as usize..to[i + 1] as usize)
                    .map(|k| (Epoch::from_offset_seconds(tt[k]), tb[k] as usize))
                    .collect(),
                apsides: (ao[i] as usize..ao[i + 1] as usize)
                    .map(|k| (Epoch::from_offset_seconds(at[k]), ad[k], ab[k] as usize, ak[k] == 1))
                    .collect(),
            })
            .collect())
    }

    /// `RelativeTrajectory::state_vector` of one ship's spline w.r.t. a body (trajectory.rs:315-335), batched over times.
    pub fn evaluate_relative(
        &mut self,
        ship: usize,
        reference: Option<usize>,
        times: &[Epoch],
    ) -> Result<Vec<Option<(DVec3, DVec3)>>, CudaPropagatorError> {
        let t: Vec<f64> = times.iter().map(|e| e.as_offset_seconds()).collect();
        let (mut p, mut v, mut ok) = (vec![0.0f64; 3 * t.len()], vec![0.0f64; 3 * t.len()], vec![0i32; t.len()]);
        status(unsafe {
            ee_ships_evaluate_relative(
                self.handle,
                ship as i64,
                reference.map_or(-1, |r| r as i32),
                t.len() as i64,
                t.as_ptr(),
                p.as_mut_ptr(),
                v.as_mut_ptr(),
                ok.as_mut_ptr(),
            )
        })?;
        Ok((0..t.len())
            .map(|k| {
                (ok[k] != 0).then(|| {
                    (DVec3::new(p[3 * k], p[3 * k + 1], p[3 * k + 2]), DVec3::new(v[3 * k], v[3 * k + 1], v[3 * k + 2]))
                })
            })
            .collect())
    }
}

impl Drop for CudaSpacecraftBatch<'_> {
    fn drop(&mut self) {
        unsafe { ee_ships_destroy(self.handle) }
    }
}
