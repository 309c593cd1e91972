// Minimal standard disciplines: the compiler only needs the discipline names to
// recognise net declarations and the access functions V()/I()/Temp()/Pwr().
`ifdef DISCIPLINES_VAMS
`else
`define DISCIPLINES_VAMS 1
discipline electrical
  potential Voltage;
  flow Current;
enddiscipline
discipline thermal
  potential Temperature;
  flow Power;
enddiscipline
`endif
