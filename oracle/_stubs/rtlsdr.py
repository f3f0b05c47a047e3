# Stub so that /root/reference/src/gpsrecv.py imports without the RTL-SDR driver
# (only used by oracle/make_golden.py in the build container).
class RtlSdr:  # pragma: no cover
    pass
