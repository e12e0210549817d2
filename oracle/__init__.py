"""CPU oracles (test infrastructure only; see the module headers).  Never imported by the
product package `denseretrievaltoolkits_b200`."""
