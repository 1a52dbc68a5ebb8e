"""pytest plugin: `python -m pytest -p refplugin /root/reference/tests` runs the reference's own tests over the
NumPy stand-ins (see jaxshim/README.md)."""
import refimport

refimport.install()
