"""Without tracing there is no runtime wrapper around a host callback's exception: the callback's own ValueError
propagates, so `JaxRuntimeError` is that class here."""

JaxRuntimeError = ValueError
