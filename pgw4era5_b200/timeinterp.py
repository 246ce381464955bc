"""
Calendar handling of the climate deltas: which two delta stamps bracket an ERA5
date and with which abscissae.  Host-side mirror of the time logic in the
reference's ``load_delta`` (functions.py:223-292); the arithmetic itself (the
two-point blend) runs on the GPU.
"""
from datetime import datetime, timezone

import numpy as np

_EPOCH = np.datetime64("1970-01-01T00:00:00", "ns")
_NS = np.timedelta64(1, "ns")


def to_datetime(stamp):
    """numpy datetime64 -> naive UTC ``datetime`` (functions.py:39-51)."""
    secs = (np.datetime64(stamp, "ns") - _EPOCH) / np.timedelta64(1, "s")
    return datetime.fromtimestamp(float(secs), tz=timezone.utc).replace(tzinfo=None)


def drop_leap_day(stamps):
    """Indices of the stamps kept after removing 29 February (functions.py:224-230)."""
    stamps = np.asarray(stamps).astype("datetime64[ns]")
    leap = None
    for i, s in enumerate(stamps):
        d = to_datetime(s)
        if d.month == 2 and d.day == 29:
            leap = i
    return [i for i in range(len(stamps)) if i != leap]


class TimeBracket:
    """Result of bracketing one ERA5 date: indices into the *kept* stamps and the
    float64 nanosecond abscissae xarray would hand to scipy's interp1d."""
    __slots__ = ("ind_before", "ind_after", "x_hi", "x_new")

    def __init__(self, ind_before, ind_after, x_hi, x_new):
        self.ind_before, self.ind_after, self.x_hi, self.x_new = ind_before, ind_after, x_hi, x_new

    @property
    def exact(self):
        return self.ind_before == self.ind_after


def moved_to_year(kept_stamps, year):
    """The stamps with their year replaced by ``year`` (functions.py:235-238)."""
    return np.array([np.datetime64(to_datetime(s).replace(year=year), "ns") for s in kept_stamps])


def bracket(kept_stamps, target, moved=None):
    """
    functions.py:233-292.  ``kept_stamps``: datetime64 stamps with 29 Feb already
    dropped; ``target``: naive ``datetime``.  Stamps are moved to the target's year;
    a target before the first (after the last) stamp wraps to the last stamp of
    the previous year (first stamp of the next year).
    """
    year = target.year
    if moved is None:                  # callers that bracket many dates pass moved_to_year(kept_stamps, year)
        moved = moved_to_year(kept_stamps, year)
    tgt = np.datetime64(target, "ns")
    before = np.nonzero(moved <= tgt)[0]
    after = np.nonzero(moved >= tgt)[0]
    if len(before):
        ib, tb = int(before[-1]), moved[before[-1]]
    else:
        ib, tb = -1, np.datetime64(to_datetime(moved[-1]).replace(year=year - 1), "ns")
    if len(after):
        ia, ta = int(after[0]), moved[after[0]]
    else:
        ia, ta = 0, np.datetime64(to_datetime(moved[0]).replace(year=year + 1), "ns")
    if ib == ia:
        return TimeBracket(ib, ia, 1.0, 0.0)
    return TimeBracket(ib, ia, float((ta - tb) / _NS), float((tgt - tb) / _NS))
