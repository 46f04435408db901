"""`integrators` of the reference: importing the four plugin modules registers nothing here -- scenes are built with
sdfrt.integrators.create(name) by the factories registered with sdfrt.compat.register_scene."""
