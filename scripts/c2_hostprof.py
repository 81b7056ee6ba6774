import sys, time
sys.path.insert(0, ".")
import ephemeris_explorer_b200 as ee
s = ee.formats.load_system("tests/golden/systems/full_solar_system_2433282.5.json")
p = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu, solout=(s.dt, s.sample_period, s.degree))
p.step(12); p.sync()
for k in (8000, 8000, 8000, 40000, 100000):
    t0 = time.perf_counter(); p.step(k); p.sync(); dt = time.perf_counter() - t0
    print("steps %d  %.1f ms  -> %.0f steps/s" % (k, dt * 1e3, k / dt))
t0 = time.perf_counter(); sol = p.take_solution(); print("take_solution %.1f ms, polys %d" % ((time.perf_counter() - t0) * 1e3, sum(len(x.polynomials) for x in sol)))
