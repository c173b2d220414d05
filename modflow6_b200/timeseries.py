"""Time series of MODFLOW 6 input decks (TS6 files of the stress packages): host-side input handling on the caller
side of the path -- the solve only ever sees the numbers of the current time step.

Follows src/Utilities/TimeSeries/TimeSeries.f90: `GetValue` :209-232 (STEPWISE / LINEAR: the time-weighted average over
the step; LINEAREND: the value at the end of the step), `get_value_at_time` :478-554, `get_integrated_value` :561-690,
`get_average_value` :697-722, and TimeSeriesManager.f90 `tsmgr_ad` :134-260 (begin = totimc, end = totimc + delt)."""
import numpy as np

DSAME = 100.0 * np.finfo(np.float64).eps     # Constants.f90:122
DNODATA = 3.0e30                             # Constants.f90: "no data" entries of a time-series file are skipped


class TimeSeriesError(ValueError):
    pass


def is_close(a, b):
    """MathUtil.f90:45-86 with the default arguments"""
    return a == b or abs(a - b) <= DSAME * max(abs(a), abs(b))


class TimeSeries:
    def __init__(self, name, method, times, values):
        self.name = name.upper()
        self.method = method.upper()
        if self.method not in ("STEPWISE", "LINEAR", "LINEAREND"):
            raise TimeSeriesError(f"time series {name}: unknown interpolation method {method}")
        self.t = np.asarray(times, dtype=np.float64)
        self.v = np.asarray(values, dtype=np.float64)
        if self.t.size == 0:
            raise TimeSeriesError(f"time series {name}: no records")
        if np.any(np.diff(self.t) <= 0.0):
            raise TimeSeriesError(f"time series {name}: times must increase")

    def value_at(self, time):
        t, v = self.t, self.v
        i = int(np.searchsorted(t, time, side="right")) - 1       # latest record with t <= time
        if i >= 0 and is_close(t[i], time):
            return float(v[i])
        if i + 1 < t.size and is_close(t[i + 1], time):
            return float(v[i + 1])
        if i < 0:
            raise TimeSeriesError(f'Error getting value at time {time:g} for time series "{self.name}"')
        if i + 1 >= t.size:                                         # only an earlier record
            if self.method == "STEPWISE":
                return float(v[i])
            raise TimeSeriesError(f'Error getting value at time {time:g} for time series "{self.name}"')
        if self.method == "STEPWISE":
            return float(v[i])
        ratio = (time - t[i]) / (t[i + 1] - t[i])
        return float(v[i] + ratio * (v[i + 1] - v[i]))

    def integrated(self, time0, time1):
        t, v = self.t, self.v
        i = int(np.searchsorted(t, time0, side="right")) - 1
        if i < 0 and is_close(t[0], time0):
            i = 0
        if i < 0:
            raise TimeSeriesError(f'Error encountered while performing integration for time series "{self.name}" '
                                  f"for time interval: {time0:g} to {time1:g}")
        value = 0.0
        while True:
            cur = t[i]
            if is_close(cur, time1) or cur > time1:
                break
            if i + 1 >= t.size:
                raise TimeSeriesError(f'Error encountered while performing integration for time series "{self.name}" '
                                      f"for time interval: {time0:g} to {time1:g}")
            nxt = t[i + 1]
            a = cur if (cur > time0 or is_close(cur, time0)) else time0
            b = nxt if (nxt < time1 or is_close(nxt, time1)) else time1
            if self.method == "STEPWISE":
                value += v[i] * (b - a)
            else:
                dv = v[i + 1] - v[i]
                v0 = v[i] + (a - cur) / (nxt - cur) * dv
                v1 = v[i] + (b - cur) / (nxt - cur) * dv
                if self.method == "LINEAR":
                    value += 0.5 * (b - a) * (v0 + v1)
                else:                       # LINEAREND: no area, the value at the end of the span
                    value = v1
            if b > time1 or is_close(b, time1):
                break
            i += 1
        return float(value)

    def value(self, time0, time1):
        """GetValue: what a boundary linked to this series uses during the time step [time0, time1]"""
        if self.method == "LINEAREND":
            return self.value_at(time1)
        if time1 - time0 > 0.0:
            return self.integrated(time0, time1) / (time1 - time0)
        return self.value_at(time0)


def read_ts_file(path, read_blocks):
    """ATTRIBUTES (NAME(S), METHOD(S), SFAC(S)) + TIMESERIES (time value1 value2 ...), utl-ts.dfn.  `read_blocks` is
    mf6io.read_blocks.  Returns {NAME: TimeSeries}."""
    names, methods, sfacs, rows = [], [], [], []
    for nm, _, lines in read_blocks(path):
        if nm == "ATTRIBUTES":
            for ln in lines:
                key = ln[0].upper()
                if key in ("NAME", "NAMES"):
                    names = [x.upper() for x in ln[1:]]
                elif key in ("METHOD", "METHODS"):
                    methods = [x.upper() for x in ln[1:]]
                elif key in ("SFAC", "SFACS"):
                    sfacs = [float(x) for x in ln[1:]]
        elif nm == "TIMESERIES":
            rows = [[float(x) for x in ln] for ln in lines]
    if not names:
        raise TimeSeriesError(f"{path}: no NAMES in the ATTRIBUTES block")
    n = len(names)
    if len(methods) == 1:
        methods = methods * n
    if len(methods) != n:
        raise TimeSeriesError(f"{path}: {n} names, {len(methods)} interpolation methods")
    if len(sfacs) == 1:
        sfacs = sfacs * n
    sfacs = sfacs or [1.0] * n
    out = {}
    for j, name in enumerate(names):
        tv = [(r[0], r[1 + j]) for r in rows if len(r) > 1 + j and r[1 + j] != DNODATA]
        out[name] = TimeSeries(name, methods[j], [a for a, _ in tv], [b * sfacs[j] for _, b in tv])
    return out
