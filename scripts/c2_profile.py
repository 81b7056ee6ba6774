import sys, os, time
sys.path.insert(0, ".")
import ephemeris_explorer_b200 as ee
s = ee.formats.load_system("tests/golden/systems/full_solar_system_2433282.5.json")
p = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu, solout=(s.dt, s.sample_period, s.degree))
p.step(12); p.sync()
os.environ["EE_SMALL_PROFILE"] = "1"
p.step(4000); p.sync()
del os.environ["EE_SMALL_PROFILE"]
t0 = time.perf_counter(); p.step(400000); p.sync(); print("C2 steps/s", 400000 / (time.perf_counter() - t0))
p2 = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu)
p2.step(12); p2.sync()
t0 = time.perf_counter(); p2.step(400000); p2.sync(); print("C2 no-solout steps/s", 400000 / (time.perf_counter() - t0))
