int dummy;
