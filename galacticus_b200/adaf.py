"""Construction-time tabulations of ``accretionDisksADAF`` (source/accretion_disks/ADAF.F90:312-451): jet power per unit
accretion rate and the spin-up function of an advection-dominated accretion flow in the model of Benson & Babul (2009),
tabulated against the inverse spin 1 - j.  The reference builds the table once in its constructor; the device only
interpolates it (GLC_TABLE_ADAF), so this restatement lives on the host, in numpy scalars.

Structure of the flow (all in gravitational units, r in GM/c^2): the fitting functions of Benson & Babul (2009) for the
radial velocity (:1053-1181), temperature (:972-1051), enthalpy x angular momentum (:856-944) and height (:1183-1241) as the
reference codes them, the Kerr-metric factors of source/black_holes/fundamentals.F90, and from them the jet powers launched
from the black hole (static limit) and from the disk (ISCO) (:548-672).

Pinned by the reference's own unit test source/tests/accretion_disks.F90:37-66 (jet power efficiency at six spins to 1e-3):
the black-hole known-answer tests under tests/ (test_adaf_jet_power_known_answers).
"""
import math

import numpy as np

SPEED_LIGHT_KMS = 2.99792458e8 / 1.0e3


# ---------------------------------------------------------------- Kerr metric (black_holes/fundamentals.F90)
def _a1(j):
    return 1.0 + ((1.0 - j * j) ** (1.0 / 3.0)) * ((1.0 + j) ** (1.0 / 3.0) + (1.0 - j) ** (1.0 / 3.0))  # :560


def _a2(j):
    return math.sqrt(3.0 * j * j + _a1(j) ** 2)  # :571


def isco_radius(j):
    """Black_Hole_ISCO_Radius_Spin, prograde (:78-121)."""
    a1, a2 = _a1(j), _a2(j)
    return 3.0 + a2 - math.sqrt((3.0 - a1) * (3.0 + a1 + 2.0 * a2))


def isco_specific_energy(j):
    """Black_Hole_ISCO_Specific_Energy_Spin (:207-233)."""
    if j >= 0.99999:
        return 0.5773502693 + 0.9164864242 * (1.0 - j) ** (1.0 / 3.0)
    r = isco_radius(j)
    return (r * r - 2.0 * r + j * math.sqrt(r)) / r / math.sqrt(r * r - 3.0 * r + 2.0 * j * math.sqrt(r))


def horizon_radius(j):
    return 1.0 + math.sqrt(1.0 - j * j)  # :487


def static_radius(j):
    return 1.0 + math.sqrt(1.0 - (j * math.cos(math.pi / 2.0)) ** 2)  # equatorial plane, :540-551


def metric_a(j, r):
    return 1.0 + j / r**2 + 2.0 * j * j / r**3  # Black_Hole_Metric_A_Factor_Spin :392 (as coded)


def metric_d(j, r):
    return 1.0 - 2.0 / r + (j / r) ** 2  # Black_Hole_Metric_D_Factor_Spin :441


def frame_dragging(j, r):
    return 2.0 * j / metric_a(j, r) / r**3  # Black_Hole_Frame_Dragging_Frequency_Spin :343


def rotational_energy_spin_down(j):
    """Black_Hole_Rotational_Energy_Spin_Down_Spin: S = [(1 + sqrt(1 - j^2))^2 + j^2] sqrt(1 - j^2) / j."""
    if j > 5.0e-8:
        return ((1.0 + math.sqrt(1.0 - j * j)) ** 2 + j * j) * math.sqrt(1.0 - j * j) / j
    if j > 1.0e-20:
        return 4.0 / j - 3.0 * j - 0.25 * j**3
    return 0.0


class ADAF:
    """accretionDisksADAF for one choice of options.  energy: 'pureADAF' | 'ISCO'; field: 'exponential' | 'linear';
    viscosity: 'fit' | a fixed alpha."""

    def __init__(self, energy="pureADAF", field="exponential", viscosity="fit", adiabatic_index=1.444, efficiency_jet_maximum=2.0):
        self.energy, self.field, self.viscosity = energy, field, viscosity
        self.gamma_ad = adiabatic_index
        self.efficiency_jet_maximum = efficiency_jet_maximum
        self.pressure_thermal_fractional = (8.0 - 6.0 * adiabatic_index) / 3.0 / (1.0 - adiabatic_index)  # :348

    # ---- adafViscosityParameter :733-763
    def alpha(self, j):
        if self.viscosity != "fit":
            return float(self.viscosity)
        if self.energy == "ISCO":
            return 0.015 + 0.02 * j**4 if self.field == "exponential" else 0.025 + 0.08 * j**4
        return 0.010 if self.field == "exponential" else 0.025 + 0.02 * j**4

    # ---- adafVelocity :1053-1181
    def velocity(self, j, r):
        rh, ri = horizon_radius(j), isco_radius(j)
        x, xh = r / ri, rh / ri
        g = self.gamma_ad - 1.444
        alpha_eff = self.alpha(j) * (1.0 + 6.450 * g + 1.355 * g * g)
        p1 = 9.0 * math.log(9.0 * x)
        p2 = math.exp(-0.66 * (1.0 - 2.0 * alpha_eff) * math.log(alpha_eff / 0.1) * math.log(x / xh))
        p3 = 1.0 - math.exp(-x * (0.16 * (j - 1.0) + 0.76))
        p4 = (1.4 + 0.29065 * (j - 0.5) ** 4 - 0.87560 * (j - 0.5) ** 2
              + (-0.33 * j + 0.45035) * (1.0 - math.exp(-(x - xh))))
        p5 = 2.3 * math.exp(40.0 * (j - 1.0)) * math.exp(-15.0 * ri * (x - xh)) + 1.0
        phi = p1 * p2 * p3 * p4 * p5
        re = rh + phi * (r - rh)
        return math.sqrt(1.0 - (1.0 - 2.0 / re + (j / re) ** 2))

    # ---- adafTemperature :972-1051
    def temperature(self, j, r):
        la = math.log10(self.alpha(j))
        ri = isco_radius(j)
        g = self.gamma_ad
        t1 = -0.270278 * g + 1.360270
        t2 = -0.9400 + 4.4744 * (g - 1.444) - 5.1402 * (g - 1.444) ** 2
        t3 = 0.840 * la + 0.919 - 0.643 * math.exp(-0.209 / self.alpha(j))
        t4 = (0.6365 * ri - 0.4828) * (1.0 + 11.9 * math.exp(-0.838 * ri**4))
        t5 = 1.444 * math.exp(-1.01 * ri**0.86) + 0.1
        return 0.31 * ((1.0 + (t4 / r) ** 0.9) ** (t2 + t3)) / (r - t5) ** t1

    def enthalpy(self, j, r):  # adafEnthalpy :946-970
        return 1.0 + (self.gamma_ad / (self.gamma_ad - 1.0)) * self.temperature(j, r)

    # ---- adafEnthalpyAngularMomentumProduct :856-944
    def enthalpy_angular_momentum(self, j, r):
        la = math.log10(self.alpha(j))
        ri = isco_radius(j)
        g = self.gamma_ad
        e1 = 0.0871 * ri - 0.10282
        e2 = 0.5000 - 7.7983 * (g - 1.333) ** 1.26
        e3 = 0.153 * (ri - 0.6) ** 0.30 + 0.105
        e4 = e3 * (0.9000 * g - 0.2996) * (1.202 - 0.080 * (la + 2.5) ** 2.6)
        e5 = -1.800 * g + 4.299 - 0.018 + 0.018 * (la + 2.0) ** 3.571
        e6 = e4 * (((0.14 * math.log10(r) ** e5 + 0.23) / e4) ** 10.0 + 1.0) ** 0.1
        return e2 + (e1 + 10.0**e6) * (1.15 - 0.03 * (3.0 + la) ** 2.37)

    def angular_momentum(self, j, r):  # adafAngularMomentum :837-854
        return self.enthalpy_angular_momentum(j, r) / self.enthalpy(j, r)

    def gamma_radial(self, j, r):  # :813-835
        return math.sqrt(1.0 / (1.0 - self.velocity(j, r) ** 2))

    def gamma_azimuthal(self, j, r):  # :784-811
        return math.sqrt(1.0 + ((self.angular_momentum(j, r) / self.gamma_radial(j, r) / r) ** 2) / metric_a(j, r))

    def gamma(self, j, r):  # :765-782
        return self.gamma_radial(j, r) * self.gamma_azimuthal(j, r)

    def fluid_angular_velocity(self, j, r):  # :707-731
        return (self.angular_momentum(j, r) * math.sqrt(metric_d(j, r) / metric_a(j, r) ** 3) / r**2
                / self.gamma_azimuthal(j, r) / self.gamma_radial(j, r))

    def field_enhancement(self, j, r):  # :674-705
        t_phi = 1.0 / self.fluid_angular_velocity(j, r)
        t_r = r * self.gamma_azimuthal(j, r) / self.velocity(j, r) / math.sqrt(metric_d(j, r))
        t = min(t_phi, t_r)
        w = frame_dragging(j, r)
        return math.exp(w * t) if self.field == "exponential" else 1.0 + w * t

    def height(self, j, r):  # adafHeight :1183-1241
        w = frame_dragging(j, r)
        L, gphi, gr = self.angular_momentum(j, r), self.gamma_azimuthal(j, r), self.gamma_radial(j, r)
        a, d = metric_a(j, r), metric_d(j, r)
        nu2 = (j * j + (1.0 - (j * w) ** 2) * L * L - ((j * gphi) ** 2 / a) * gr * gr * d
               - gr * math.sqrt(d / a) * 2.0 * L * w * gphi * j * j) / r**4
        return math.sqrt(self.temperature(j, r) / self.enthalpy(j, r) / r**2 / nu2)

    def _jet_common(self, j, r):
        beta_phi = math.sqrt(1.0 - 1.0 / self.gamma_azimuthal(j, r) ** 2)
        d = metric_d(j, r)
        return ((3.0 / 80.0) * r * r * (2.0 * j * beta_phi / r**2 + math.sqrt(d)) ** 2
                * (1.0 - self.pressure_thermal_fractional) * (self.field_enhancement(j, r) * self.gamma(j, r)) ** 2
                * math.sqrt((1.0 - self.velocity(j, r) ** 2) / d))

    def jet_power_disk(self, j, r):  # adafJetPowerDisk :548-608
        return (self._jet_common(j, r) * (self.fluid_angular_velocity(j, r) + frame_dragging(j, r)) ** 2
                * self.temperature(j, r) / metric_a(j, r) / self.velocity(j, r) / self.height(j, r))

    def jet_power_black_hole(self, j, r):  # adafJetPowerBlackHole :610-672
        if not j > 5.0e-8:
            return 0.0
        return (self._jet_common(j, r) * frame_dragging(j, r) ** 2
                * self.temperature(j, r) / metric_a(j, r) / self.velocity(j, r) / self.height(j, r))

    def jet_power_disk_from_black_hole(self, j, r):  # :525-546
        return self.jet_power_disk(j, r) * (1.0 - 1.0 / self.field_enhancement(j, r) ** 2)

    # ---- what the constructor tabulates (:403-447), per unit accretion rate
    def jet_efficiency(self, j):
        return min(self.jet_power_black_hole(j, static_radius(j)) + self.jet_power_disk(j, isco_radius(j)), self.efficiency_jet_maximum)

    def spin_up(self, j):
        energy = 1.0 if self.energy == "pureADAF" else isco_specific_energy(j)
        ri = isco_radius(j)
        return (self.angular_momentum(j, ri) - 2.0 * j * energy
                - rotational_energy_spin_down(j) * (self.jet_power_black_hole(j, static_radius(j)) + self.jet_power_disk_from_black_hole(j, ri)))


def adaf_tabulations(count=10000, x_min=1.0e-6, x_max=1.0, **options):
    """GLC_TABLE_ADAF as the reference's constructor fills it: x = 1 - j on a logarithmic lattice of `count` points over
    [1e-6, 1] (table1DLogarithmicLinear, :396-402), values[:, 0] = jet power per unit accretion rate in (km/s)^2,
    values[:, 1] = spin-up to mass-rate ratio."""
    disk = ADAF(**options)
    x = np.exp(np.linspace(np.log(x_min), np.log(x_max), count))
    x[0], x[-1] = x_min, x_max
    values = np.empty((count, 2))
    for i, xi in enumerate(x):
        j = 1.0 - float(xi)
        values[i, 0] = disk.jet_efficiency(j) * SPEED_LIGHT_KMS**2
        values[i, 1] = disk.spin_up(j)
    return x, None, values
