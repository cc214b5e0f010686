/*
 * mpirun.c — launcher for minimpi jobs:  mpirun -np N [-x VAR=VAL] program [args...]
 * Forks N copies of the program with MINIMPI_RANK / MINIMPI_SIZE / MINIMPI_JOB set, waits for
 * all of them; if one dies the rest are killed.  TEST / BASELINE INFRASTRUCTURE (see mpi.h).
 */
#define _GNU_SOURCE
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

int main(int argc, char **argv)
{
   int np = 1, i = 1, r, status, worst = 0, alive;
   pid_t *pids;
   char job[64];
   while (i < argc && argv[i][0] == '-')
   {
      if ((!strcmp(argv[i], "-np") || !strcmp(argv[i], "-n")) && i + 1 < argc) { np = atoi(argv[i + 1]); i += 2; }
      else if (!strcmp(argv[i], "-x") && i + 1 < argc) { putenv(argv[i + 1]); i += 2; }
      else break;
   }
   if (i >= argc || np < 1) { fprintf(stderr, "usage: mpirun -np N program [args...]\n"); return 2; }
   snprintf(job, sizeof(job), "j%d_%ld", (int) getpid(), (long) time(NULL));
   pids = (pid_t *) calloc((size_t) np, sizeof(pid_t));
   for (r = 0; r < np; r++)
   {
      pid_t p = fork();
      if (p < 0) { perror("fork"); return 2; }
      if (p == 0)
      {
         char buf[64];
         snprintf(buf, sizeof(buf), "%d", r);  setenv("MINIMPI_RANK", buf, 1);
         snprintf(buf, sizeof(buf), "%d", np); setenv("MINIMPI_SIZE", buf, 1);
         setenv("MINIMPI_JOB", job, 1);
         execvp(argv[i], &argv[i]);
         perror("execvp");
         _exit(127);
      }
      pids[r] = p;
   }
   alive = np;
   while (alive > 0)
   {
      pid_t p = wait(&status);
      int code;
      if (p < 0) break;
      alive--;
      code = WIFEXITED(status) ? WEXITSTATUS(status) : 128 + WTERMSIG(status);
      if (code != 0)
      {
         if (!worst) worst = code;
         for (r = 0; r < np; r++) if (pids[r] != p) kill(pids[r], SIGTERM);
      }
   }
   {
      char name[128];
      snprintf(name, sizeof(name), "/dev/shm/minimpi_%s_%d", job, (int) getuid());
      unlink(name);
   }
   return worst;
}
