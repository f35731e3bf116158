"""Feeder topology generators (input data for the batched step).

Each class yields the same ``buses`` / ``lines`` / ``loads`` / ``generators``
content, in the same order, as the reference class of the same name, so bus
and line indices ("topology ordering") agree bit for bit:

* ``SimpleRadialFeeder`` - reference ``grid_fed_rl/feeders/base.py:256-303``
* ``IEEE13Bus``          - reference ``grid_fed_rl/feeders/ieee_feeders.py:22-142``
* ``IEEE34Bus``          - reference ``grid_fed_rl/feeders/ieee_feeders.py:145-233``
* ``IEEE123Bus``         - reference ``grid_fed_rl/feeders/ieee_feeders.py:236-378``
* ``SyntheticFeeder`` / ``ScalableFeeder`` - reference ``grid_fed_rl/feeders/synthetic.py:23-252``
* ``CustomFeeder``       - reference ``grid_fed_rl/feeders/base.py:160-253``

Difference by design: the reference draws IEEE-34 / IEEE-123 impedances and
loads from the *unseeded* global ``np.random`` stream, so two
constructions never agree.  Here every generator owns a
``numpy.random.RandomState(seed)`` - the same MT19937 stream the reference sees
after ``np.random.seed(seed)`` - so ``IEEE34Bus(seed=0)`` equals the reference
constructed right after ``np.random.seed(0)`` (deviation D4-i in DESIGN.md) and
the global stream is left untouched.  ``tests/test_feeders.py`` checks equality
against the reference when it is importable and against frozen fixtures
otherwise.
"""

from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .components import Bus, FeederParameters, Line, Load

FT_PER_MILE = 5280.0


class BaseFeeder:
    """Container: ordered component lists plus the per-unit bases."""

    def __init__(self, name: str, parameters: Optional[FeederParameters] = None) -> None:
        self.name = name
        self.parameters = parameters or FeederParameters(base_voltage=12.47, base_power=10.0,
                                                          frequency=60.0)
        self.buses: List[Bus] = []
        self.lines: List[Line] = []
        self.loads: List[Load] = []
        self.generators: Dict[str, Dict[str, Any]] = {}

    # -- small conveniences the reference exposes (base.py:53-91) -----------
    def add_bus(self, bus: Bus) -> None:
        self.buses.append(bus)

    def add_line(self, line: Line) -> None:
        self.lines.append(line)

    def add_load(self, load: Load) -> None:
        self.loads.append(load)

    def add_generator(self, gen_id: str, gen_info: Dict[str, Any]) -> None:
        self.generators[gen_id] = gen_info

    def get_bus_by_id(self, bus_id):
        return next((b for b in self.buses if b.id == bus_id), None)

    def get_line_by_id(self, line_id):
        return next((l for l in self.lines if l.id == line_id), None)

    def get_load_by_id(self, load_id):
        return next((l for l in self.loads if l.id == load_id), None)

    @property
    def base_impedance(self) -> float:
        return (self.parameters.base_voltage ** 2) / self.parameters.base_power

    def get_network_stats(self) -> Dict[str, Any]:
        return {
            "name": self.name,
            "num_buses": len(self.buses),
            "num_lines": len(self.lines),
            "num_loads": len(self.loads),
            "num_generators": len(self.generators),
            "total_load": sum(l.base_power for l in self.loads),
            "total_generation_capacity": sum(g.get("capacity", 0) for g in self.generators.values()),
            "base_voltage_kv": self.parameters.base_voltage,
            "base_power_mva": self.parameters.base_power,
        }

    def validate_network(self) -> List[str]:
        errors: List[str] = []
        ids = {b.id for b in self.buses}
        touched = {l.from_bus for l in self.lines} | {l.to_bus for l in self.lines}
        if len(self.buses) > 1:
            errors += [f"Bus {b.id} is not connected to any line" for b in self.buses
                       if b.id not in touched]
        for l in self.lines:
            for end in (l.from_bus, l.to_bus):
                if end not in ids:
                    errors.append(f"Line {l.id} references non-existent bus {end}")
        errors += [f"Load {l.id} references non-existent bus {l.bus}" for l in self.loads
                   if l.bus not in ids]
        n_slack = sum(b.bus_type == "slack" for b in self.buses)
        if n_slack != 1:
            errors.append("No slack bus found - at least one bus must be slack type" if n_slack == 0
                          else f"Multiple slack buses found: "
                               f"{[b.id for b in self.buses if b.bus_type == 'slack']}")
        return errors


class CustomFeeder(BaseFeeder):
    """User-assembled network; ``from_dict`` takes the reference's dict schema."""

    def __init__(self, name: str = "Custom", parameters: Optional[FeederParameters] = None) -> None:
        super().__init__(name, parameters)

    def build_network(self) -> None:  # components are added by the caller
        return None

    def from_dict(self, spec: Dict[str, Any]) -> None:
        self.buses, self.lines, self.loads, self.generators = [], [], [], {}
        default_level = self.parameters.base_voltage * 1000
        for b in spec.get("buses", []):
            self.add_bus(Bus(b["id"], b.get("voltage_level", default_level), b.get("type", "pq"),
                             b.get("base_voltage", 1.0)))
        for l in spec.get("lines", []):
            self.add_line(Line(l["id"], l["from_bus"], l["to_bus"], l["resistance"], l["reactance"],
                               l.get("rating", 1e6)))
        for d in spec.get("loads", []):
            self.add_load(Load(d["id"], d["bus"], d["power"], d.get("power_factor", 0.95)))
        for g in spec.get("generators", []):
            self.add_generator(g["id"], g)

    def to_dict(self) -> Dict[str, Any]:
        p = self.parameters
        return {
            "name": self.name,
            "parameters": {"base_voltage": p.base_voltage, "base_power": p.base_power,
                           "frequency": p.frequency},
            "buses": [{"id": b.id, "voltage_level": b.voltage_level, "type": b.bus_type,
                       "base_voltage": b.base_voltage} for b in self.buses],
            "lines": [{"id": l.id, "from_bus": l.from_bus, "to_bus": l.to_bus,
                       "resistance": l.resistance, "reactance": l.reactance, "rating": l.rating}
                      for l in self.lines],
            "loads": [{"id": d.id, "bus": d.bus, "power": d.base_power,
                       "power_factor": d.power_factor} for d in self.loads],
            "generators": list(self.generators.values()),
        }


class SimpleRadialFeeder(BaseFeeder):
    """Chain 1-2-...-N, slack at bus 1, one load on every other bus."""

    def __init__(self, num_buses: int = 5, line_impedance: Tuple[float, float] = (0.01, 0.02),
                 load_power: float = 1e6, name: str = "SimpleRadial") -> None:
        super().__init__(name)
        self.num_buses = num_buses
        self.line_impedance = line_impedance
        self.load_power = load_power
        self.build_network()

    def build_network(self) -> None:
        level = self.parameters.base_voltage * 1000
        r, x = self.line_impedance
        for k in range(1, self.num_buses + 1):
            self.add_bus(Bus(k, level, "slack" if k == 1 else "pq"))
        for k in range(1, self.num_buses):
            self.add_line(Line(f"line_{k}_{k + 1}", k, k + 1, r, x, 5e6))
        for k in range(2, self.num_buses + 1):
            self.add_load(Load(f"load_{k}", k, self.load_power, 0.95))


# --------------------------------------------------------------------------- IEEE 13

_IEEE13_BUSES = (650, 632, 633, 634, 645, 646, 671, 680, 684, 611, 652, 692, 675)
# (from, to, length in feet, conductor configuration)
_IEEE13_SECTIONS = (
    (650, 632, 2000, "601"), (632, 633, 500, "602"), (632, 645, 500, "603"),
    (632, 671, 2000, "601"), (645, 646, 300, "603"), (671, 680, 1000, "601"),
    (671, 684, 300, "604"), (633, 634, 0, "XFM1"), (684, 611, 300, "603"),
    (684, 652, 800, "607"), (671, 692, 0, "SWITCH"), (692, 675, 500, "606"),
)
# ohm/mile (R, X) per configuration; XFM1 / SWITCH have zero length in the table above
IEEE13_CONFIG_Z = {
    "601": (0.3465, 1.0179), "602": (0.7526, 1.1814), "603": (1.3238, 1.3569),
    "604": (1.3238, 1.3569), "606": (0.7982, 0.4463), "607": (1.3425, 0.5124),
    "XFM1": (0.0, 0.06), "SWITCH": (0.0001, 0.0001),
}
# (bus, kW, kVAr)
_IEEE13_SPOT_LOADS = ((634, 400, 290), (645, 170, 125), (646, 230, 132), (652, 128, 86),
                      (671, 1155, 660), (675, 843, 462), (692, 170, 151), (611, 170, 80))


class IEEE13Bus(BaseFeeder):
    def __init__(self) -> None:
        super().__init__("IEEE13Bus", FeederParameters(base_voltage=4.16, base_power=10.0,
                                                       frequency=60.0))
        self.build_network()

    def build_network(self) -> None:
        level = self.parameters.base_voltage * 1000
        for bid in _IEEE13_BUSES:
            self.add_bus(Bus(bid, level, "slack" if bid == 650 else "pq", 1.0))
        zb = self.base_impedance
        for a, b, feet, cfg in _IEEE13_SECTIONS:
            r_mile, x_mile = IEEE13_CONFIG_Z[cfg]
            miles = feet / FT_PER_MILE
            self.add_line(Line(f"line_{a}_{b}", a, b, (r_mile * miles) / zb, (x_mile * miles) / zb,
                               5e6))
        for bid, kw, kvar in _IEEE13_SPOT_LOADS:
            pf = kw / np.sqrt(kw ** 2 + kvar ** 2) if kvar != 0 else 0.95
            self.add_load(Load(f"load_{bid}", bid, (kw / 1000.0) * 1e6, pf))
        self.add_generator("solar_671", {"type": "solar", "bus": 671, "capacity": 500e3,
                                         "efficiency": 0.18})
        self.add_generator("wind_675", {"type": "wind", "bus": 675, "capacity": 1e6,
                                        "cut_in_speed": 3.0, "rated_speed": 12.0,
                                        "cut_out_speed": 25.0})


# --------------------------------------------------------------------------- IEEE 34

_IEEE34_BUSES = (800, 802, 806, 808, 810, 812, 814, 850, 816, 818, 820, 822, 824, 826, 828, 830,
                 854, 856, 858, 864, 834, 860, 836, 840, 842, 844, 846, 848, 832, 888, 890, 838,
                 862, 868)
_IEEE34_SECTIONS = (
    (800, 802), (802, 806), (806, 808), (808, 810), (810, 812), (812, 814), (814, 850),
    (816, 818), (816, 824), (818, 820), (820, 822), (824, 826), (824, 828), (828, 830),
    (854, 856), (832, 858), (858, 864), (858, 834), (834, 860), (860, 836), (836, 840),
    (840, 842), (842, 844), (844, 846), (846, 848), (832, 888), (888, 890), (890, 838),
    (834, 862), (862, 838), (842, 868),
)
_IEEE34_LOAD_BUSES = (806, 810, 820, 822, 826, 830, 854, 858, 864, 840, 844, 848, 890)


class IEEE34Bus(BaseFeeder):
    """``seed`` pins the stream the reference leaves unseeded (D4-i)."""

    def __init__(self, seed: Optional[int] = 0) -> None:
        super().__init__("IEEE34Bus", FeederParameters(base_voltage=24.9, base_power=10.0,
                                                       frequency=60.0))
        self.seed = seed
        self._rs = np.random.RandomState(seed) if seed is not None else np.random
        self.build_network()

    def build_network(self) -> None:
        rs = self._rs
        level = self.parameters.base_voltage * 1000
        for bid in _IEEE34_BUSES:
            self.add_bus(Bus(bid, level, "slack" if bid == 800 else "pq"))
        for a, b in _IEEE34_SECTIONS:
            r = 0.005 + 0.002 * rs.random_sample()
            x = 0.01 + 0.005 * rs.random_sample()
            self.add_line(Line(f"line_{a}_{b}", a, b, r, x, 10e6))
        for bid in _IEEE34_LOAD_BUSES:
            kw = 100 + 400 * rs.random_sample()
            self.add_load(Load(f"load_{bid}", bid, kw * 1000, 0.95))
        self.add_generator("solar_farm_830", {"type": "solar", "bus": 830, "capacity": 2e6,
                                              "efficiency": 0.20})


# --------------------------------------------------------------------------- IEEE 123

_IEEE123_BACKBONE = (1, 3, 7, 13, 18, 25, 35, 49, 64, 78, 97, 114)


class IEEE123Bus(BaseFeeder):
    """Backbone impedances come from ``RandomState(seed)`` (the reference draws
    them before its own ``np.random.seed(42)``); laterals/secondaries from the
    stream seeded 42 and loads/DG/storage from the stream seeded 123, exactly
    where the reference reseeds (ieee_feeders.py:285, :331)."""

    def __init__(self, seed: Optional[int] = 0) -> None:
        super().__init__("IEEE123Bus", FeederParameters(base_voltage=4.16, base_power=10.0,
                                                        frequency=60.0))
        self.seed = seed
        self._rs = np.random.RandomState(seed) if seed is not None else np.random
        self.build_network()

    def build_network(self) -> None:
        level = self.parameters.base_voltage * 1000
        bus_ids = list(range(1, 124))
        for bid in bus_ids:
            self.add_bus(Bus(bid, level, "slack" if bid == 1 else "pq"))

        rs = self._rs
        for a, b in zip(_IEEE123_BACKBONE[:-1], _IEEE123_BACKBONE[1:]):
            r = 0.003 + 0.002 * rs.random_sample()
            x = 0.006 + 0.004 * rs.random_sample()
            self.add_line(Line(f"main_{a}_{b}", a, b, r, x, 15e6))

        rs = np.random.RandomState(42)
        laterals: List[Tuple[int, int]] = []
        for trunk in _IEEE123_BACKBONE[1:]:
            k = rs.randint(2, 6)
            pool = [b for b in bus_ids if b not in _IEEE123_BACKBONE and b > trunk]
            if len(pool) >= k:
                laterals += [(trunk, tip) for tip in rs.choice(pool, k, replace=False)]
        for a, b in laterals:
            r = 0.008 + 0.005 * rs.random_sample()
            x = 0.012 + 0.008 * rs.random_sample()
            self.add_line(Line(f"lateral_{a}_{b}", a, b, r, x, 5e6))

        ties: List[Tuple[int, int]] = []
        tips = set(tip for _, tip in laterals)
        for _ in range(20):
            pool = list(tips)
            if len(pool) >= 2:
                a, b = rs.choice(pool, 2, replace=False)
                if a != b:
                    ties.append((a, b))
        for a, b in ties:
            r = 0.010 + 0.008 * rs.random_sample()
            x = 0.015 + 0.010 * rs.random_sample()
            self.add_line(Line(f"secondary_{a}_{b}", a, b, r, x, 3e6))

        rs = np.random.RandomState(123)
        for bid in bus_ids[1:]:
            if rs.random_sample() < 0.7:
                kw = 10 + 190 * rs.random_sample()
                pf = 0.92 + 0.06 * rs.random_sample()
                self.add_load(Load(f"load_{bid}", bid, kw * 1000, pf))

        for k, bid in enumerate((25, 49, 78, 97, 114)):
            if k % 2 == 0:
                cap = (200 + 300 * rs.random_sample()) * 1000
                eff = 0.18 + 0.04 * rs.random_sample()
                self.add_generator(f"solar_{bid}", {"type": "solar", "bus": bid, "capacity": cap,
                                                    "efficiency": eff})
            else:
                cap = (500 + 1000 * rs.random_sample()) * 1000
                self.add_generator(f"wind_{bid}", {"type": "wind", "bus": bid, "capacity": cap,
                                                   "cut_in_speed": 3.0, "rated_speed": 12.0,
                                                   "cut_out_speed": 25.0})
        for bid in (35, 64, 97):
            kwh = 500 + 500 * rs.random_sample()
            kw = 250 + 250 * rs.random_sample()
            eff = 0.90 + 0.05 * rs.random_sample()
            self.add_generator(f"battery_{bid}", {"type": "battery", "bus": bid,
                                                  "capacity_kwh": kwh, "power_rating_kw": kw,
                                                  "efficiency": eff})


# --------------------------------------------------------------------------- synthetic

class NetworkConfig:
    def __init__(self, num_buses: int = 20, connectivity: float = 0.3,
                 load_probability: float = 0.7, min_load_kw: float = 50, max_load_kw: float = 500,
                 line_length_range: Tuple[float, float] = (0.1, 2.0),
                 dg_probability: float = 0.2) -> None:
        self.num_buses = num_buses
        self.connectivity = connectivity
        self.load_probability = load_probability
        self.min_load_kw = min_load_kw
        self.max_load_kw = max_load_kw
        self.line_length_range = line_length_range
        self.dg_probability = dg_probability


class SyntheticFeeder(BaseFeeder):
    """Random spanning tree (+ optional extra ties) with random loads and DG.

    Stream consumption follows reference ``synthetic.py:75-225`` call for call,
    including ``choice`` over ``list(set)`` iteration order, so a given
    ``(config, seed)`` reproduces the reference network.  ``connectivity=0``
    yields a radial tree."""

    def __init__(self, config: Optional[NetworkConfig] = None, seed: Optional[int] = None,
                 name: str = "Synthetic") -> None:
        super().__init__(name)
        self.config = config or NetworkConfig()
        self.seed = seed
        self._rs = np.random.RandomState(seed) if seed is not None else np.random
        self.build_network()

    def build_network(self) -> None:
        cfg, rs = self.config, self._rs
        level = self.parameters.base_voltage * 1000
        for k in range(cfg.num_buses):
            self.add_bus(Bus(k + 1, level, "slack" if k == 0 else "pq"))

        # spanning tree grown from the slack bus
        reached = {1}
        pending = set(range(2, cfg.num_buses + 1))
        serial = 1
        while pending:
            a = rs.choice(list(reached))
            b = rs.choice(list(pending))
            self.add_line(self._random_line(serial, a, b))
            reached.add(b)
            pending.remove(b)
            serial += 1

        # extra ties
        n = cfg.num_buses
        have = len(self.lines)
        want = int(have + cfg.connectivity * (n * (n - 1) // 2 - have))
        serial = have + 1
        pairs = {(l.from_bus, l.to_bus) for l in self.lines}
        pairs |= {(l.to_bus, l.from_bus) for l in self.lines}
        tries = 0
        while len(self.lines) < want and tries < 1000:
            a = rs.randint(1, n + 1)
            b = rs.randint(1, n + 1)
            if a != b and (a, b) not in pairs:
                self.add_line(self._random_line(serial, a, b))
                pairs.add((a, b))
                pairs.add((b, a))
                serial += 1
            tries += 1

        for bus in self.buses[1:]:
            if rs.random_sample() < cfg.load_probability:
                kw = cfg.min_load_kw + (cfg.max_load_kw - cfg.min_load_kw) * rs.random_sample()
                pf = 0.85 + 0.15 * rs.random_sample()
                self.add_load(Load(f"load_{bus.id}", bus.id, kw * 1000, pf))

        self._add_distributed_generation()

    def _random_line(self, serial: int, a, b) -> Line:
        rs = self._rs
        lo, hi = self.config.line_length_range
        km = lo + (hi - lo) * rs.random_sample()
        r_km = 0.2 + 0.3 * rs.random_sample()
        x_km = 0.3 + 0.4 * rs.random_sample()
        zb = self.base_impedance
        mva = 2 + 8 * rs.random_sample()
        return Line(f"line_{serial}", a, b, (r_km * km) / zb, (x_km * km) / zb, mva * 1e6)

    def _add_distributed_generation(self) -> None:
        rs = self._rs
        for bid in [b.id for b in self.buses[1:]]:
            if rs.random_sample() >= self.config.dg_probability:
                continue
            kind = rs.choice(["solar", "wind", "battery"])
            if kind == "solar":
                kw = 100 + 400 * rs.random_sample()
                self.add_generator(f"solar_{bid}", {"type": "solar", "bus": bid,
                                                    "capacity": kw * 1000,
                                                    "efficiency": 0.15 + 0.10 * rs.random_sample()})
            elif kind == "wind":
                kw = 500 + 1500 * rs.random_sample()
                self.add_generator(f"wind_{bid}", {"type": "wind", "bus": bid,
                                                   "capacity": kw * 1000,
                                                   "cut_in_speed": 2.5 + 1.0 * rs.random_sample(),
                                                   "rated_speed": 10 + 5 * rs.random_sample(),
                                                   "cut_out_speed": 20 + 10 * rs.random_sample()})
            else:
                kwh = 200 + 800 * rs.random_sample()
                self.add_generator(f"battery_{bid}", {"type": "battery", "bus": bid,
                                                      "capacity_kwh": kwh,
                                                      "power_rating_kw": kwh * 0.5,
                                                      "efficiency": 0.85 + 0.10 * rs.random_sample()})


class ScalableFeeder(SyntheticFeeder):
    """Size-scaled parameters of reference ``synthetic.py:233-252``;
    ``radial=True`` forces ``connectivity=0`` (BASELINE config 5)."""

    def __init__(self, num_buses: int, seed: Optional[int] = None, radial: bool = False) -> None:
        cfg = NetworkConfig(
            num_buses=num_buses,
            connectivity=0.0 if radial else max(0.1, min(0.6, 20.0 / num_buses)),
            load_probability=min(0.9, 0.5 + 0.01 * num_buses),
            dg_probability=min(0.4, 0.1 + 0.005 * num_buses),
            min_load_kw=20, max_load_kw=300, line_length_range=(0.05, 1.5))
        super().__init__(cfg, seed, f"Scalable{num_buses}")
